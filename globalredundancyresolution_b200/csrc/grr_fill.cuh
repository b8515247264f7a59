// grr_fill.cuh -- host: HostChain (double, as given through the C ABI) -> ChainDev<Real,NJ> kernel parameter.
#pragma once
#include "grr_device.cuh"
#include "grr_host.h"
#include <math.h>

namespace grr {

template <typename Real, int NJ>
void fill_chain(const HostChain &h, ChainDev<Real, NJ> &d) {
  memset(&d, 0, sizeof(d));
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
      d.jc[j].pad = 0;
    }
    const bool unbounded = (h.qmax[j] - h.qmin[j]) >= 2.0 * 3.14159265358979323846;
    d.bmin[j] = unbounded ? -std::numeric_limits<Real>::infinity() : (Real)h.qmin[j];
    d.bmax[j] = unbounded ? std::numeric_limits<Real>::infinity() : (Real)h.qmax[j];
    float lo = (float)h.qmin[j], hi = (float)h.qmax[j];
    if (unbounded) { lo = -3.14159265358979f; hi = 3.14159265358979f; }
    d.slo[j] = lo; d.srange[j] = hi - lo;
    d.jc[j].bmin = d.bmin[j]; d.jc[j].bmax = d.bmax[j];
    for (int k = 0; k < 3; k++) d.jc[j].tp[k] = d.tp[j][k];
  }
  d.small_angles = 1;
  d.axes_positive = 1;
  for (int j = 0; j < h.n && j < NJ; j++) {
    if (!(h.qmin[j] >= -0.785 && h.qmax[j] <= 0.785)) d.small_angles = 0;
    if (d.jc[j].kind != 0 && !(d.jc[j].sgn > 0)) d.axes_positive = 0;
  }
  for (int k = 0; k < 3; k++) d.ee[k] = (Real)h.ee_local[k];
  for (int k = 0; k < 9; k++) { d.Rtail[k] = (Real)h.R_tail[k]; d.Rgoal[k] = (Real)h.R_goal[k]; }
}


}  // namespace grr
