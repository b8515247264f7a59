// grr_fill.cuh -- host: HostChain (double, as given through the C ABI) -> ChainDev<Real,NJ> kernel parameter.
#pragma once
#include "grr_device.cuh"
#include "grr_host.h"
#include <math.h>

namespace grr {

// Margins within which a decision of the single-precision edge phase counts as a near-tie.  u = 2^-24 * scale;
// the multipliers are the measured sizes of the single- vs double-precision differences of each quantity
// (tools/flag_study.py runs the same lane machine in both precisions on the host), `scale` is the safety factor.
//   FL_F     max|F| against tol, absolute: position rows carry ~u * reach * sqrt(n) of round-off, rotation rows ~u * pi * sqrt(n)
//   FL_X     convergence-on-x test value against tolx, absolute
//   FL_ALAM  step length against alamin, relative
//   FL_REL   other comparisons (|p| vs stpmax, descent test, discriminant), relative
//   FL_LEAF  consecutive-config distance against the discontinuity threshold, relative
//   FL_ND    dist / eps against the next integer (Ndivs = ceil), absolute
//   FL_DD    relative error allowed for the directional derivative that resolves Armijo near-ties (lane_step)
//   FL_CURV  bound of |sum_r F_r d2 F_r| / |F| per unit |p|^2 (reach * n) for the second-order term of that model
//   FL_MAXIT solves with more Newton iterations than this are re-tested (the iterates of a long solve drift apart)
inline void flag_margins(HostChain &h, double scale) {
  double reach = 0;
  for (int j = 0; j < h.n; j++) reach += sqrt(h.tp[3 * j] * h.tp[3 * j] + h.tp[3 * j + 1] * h.tp[3 * j + 1] + h.tp[3 * j + 2] * h.tp[3 * j + 2]);
  reach += sqrt(h.ee_local[0] * h.ee_local[0] + h.ee_local[1] * h.ee_local[1] + h.ee_local[2] * h.ee_local[2]);
  if (h.mode != GRR_FREE && reach < 3.15) reach = 3.15;
  if (reach < 1.0) reach = 1.0;
  const double u = scale * 5.9604644775390625e-08;
  h.fl[FL_F] = u * reach * sqrt((double)h.n);
  h.fl[FL_X] = u * 4.0;
  h.fl[FL_ALAM] = u * 2048.0;
  h.fl[FL_REL] = u * 256.0;
  h.fl[FL_LEAF] = u * 8.0;
  h.fl[FL_ND] = u * 64.0;
  h.fl[FL_MAXIT] = 1e9;
  h.fl[FL_DD] = u * 2048.0;
  h.fl[FL_CURV] = reach * h.n;
}

template <typename Real, int NJ>
void fill_chain(const HostChain &h, ChainDev<Real, NJ> &d) {
  memset(&d, 0, sizeof(d));
  d.fl_f = (Real)h.fl[FL_F]; d.fl_f2 = (Real)(4.0 * h.fl[FL_F] * h.fl[FL_F]); d.fl_x = (Real)h.fl[FL_X];
  d.fl_alam = (Real)h.fl[FL_ALAM]; d.fl_rel = (Real)h.fl[FL_REL]; d.fl_leaf = (Real)h.fl[FL_LEAF]; d.fl_nd = (Real)h.fl[FL_ND];
  d.fl_dd = (Real)h.fl[FL_DD]; d.fl_curv = (Real)h.fl[FL_CURV]; d.fl_curv1 = (Real)(h.fl[FL_CURV] / h.n); d.fl_jac = (Real)(h.fl[FL_CURV] / h.n);   // |d max|F| / dx| <= reach (rotation rows: 1)
  d.fl_maxit = h.fl[FL_MAXIT] > 1e6 ? 1 << 20 : (int)h.fl[FL_MAXIT];
  d.n = h.n; d.mode = h.mode; d.nlinks_eps = h.nlinks_eps; d.max_iters = h.max_iters;
  d.tol = (Real)h.tol; d.lambda = (Real)h.lambda;
  for (int j = 0; j < h.n && j < NJ; j++) {
    const double *R = h.Rp + 9 * j, *a = h.axis + 3 * j;
    const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double K2[9], RK[9], RK2[9];
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) K2[3 * r + q] = K[3 * r] * K[q] + K[3 * r + 1] * K[3 + q] + K[3 * r + 2] * K[6 + q];
    for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) {
      RK[3 * r + q] = R[3 * r] * K[q] + R[3 * r + 1] * K[3 + q] + R[3 * r + 2] * K[6 + q];
      RK2[3 * r + q] = R[3 * r] * K2[q] + R[3 * r + 1] * K2[3 + q] + R[3 * r + 2] * K2[6 + q];
    }
    for (int k = 0; k < 9; k++) { d.A[j][k] = (Real)(R[k] + RK2[k]); d.B[j][k] = (Real)RK[k]; d.C[j][k] = (Real)RK2[k]; }
    for (int k = 0; k < 3; k++) { d.tp[j][k] = (Real)h.tp[3 * j + k]; d.ax[j][k] = (Real)a[k]; }
    {
      bool ident = true;
      for (int k = 0; k < 9; k++) ident = ident && (fabs(R[k] - ((k % 4 == 0) ? 1.0 : 0.0)) < 1e-14);
      int kind = 0;
      for (int k = 0; k < 3; k++)
        if (ident && fabs(fabs(a[k]) - 1.0) < 1e-14 && fabs(a[(k + 1) % 3]) < 1e-14 && fabs(a[(k + 2) % 3]) < 1e-14) kind = k + 1;
      d.jc[j].kind = kind;
      d.jc[j].sgn = (Real)(kind ? (a[kind - 1] > 0 ? 1.0 : -1.0) : 0.0);
      d.jc[j].spin = h.spin[j] ? 1 : 0;
    }
    // 'spin' joints have no limits in the solver; their random starts are uniform in [0, 2 pi) ([recalled] Klamp't RobotCSpace)
    const bool unbounded = h.spin[j] || (h.qmax[j] - h.qmin[j]) >= 2.0 * 3.14159265358979323846;
    d.bmin[j] = unbounded ? -std::numeric_limits<Real>::infinity() : (Real)h.qmin[j];
    d.bmax[j] = unbounded ? std::numeric_limits<Real>::infinity() : (Real)h.qmax[j];
    float lo = (float)h.qmin[j], hi = (float)h.qmax[j];
    if (unbounded) { lo = -3.14159265358979f; hi = 3.14159265358979f; }
    if (h.spin[j]) { lo = 0.0f; hi = 6.28318530717959f; }
    d.slo[j] = lo; d.srange[j] = hi - lo;
    d.jc[j].bmin = d.bmin[j]; d.jc[j].bmax = d.bmax[j];
    for (int k = 0; k < 3; k++) d.jc[j].tp[k] = d.tp[j][k];
  }
  d.spin_mask = spin_mask_of(h);
  d.small_angles = 1;
  d.axes_positive = 1;
  for (int j = 0; j < h.n && j < NJ; j++) {
    if (!(h.qmin[j] >= -0.785 && h.qmax[j] <= 0.785) || h.spin[j]) d.small_angles = 0;
    if (d.jc[j].kind != 0 && !(d.jc[j].sgn > 0)) d.axes_positive = 0;
  }
  for (int k = 0; k < 3; k++) d.ee[k] = (Real)h.ee_local[k];
  for (int k = 0; k < 9; k++) { d.Rtail[k] = (Real)h.R_tail[k]; d.Rgoal[k] = (Real)h.R_goal[k]; }
}


}  // namespace grr
