// grr_lane.cuh -- the per-lane IK machine used by every kernel of the path.
//
// Design (B200-first; measured motivation in profiles/r01_edge_v1_ncu.md):
//   * one lane = one IK instance at a time, advanced one "trip" (= one backward chain pass + the
//     Newton / line-search transition) per loop iteration.  Lanes of a warp are never in lock step on
//     an instance: a lane that finishes fetches its next instance immediately, so the only divergent
//     code is the short transition logic, never the chain pass (v1: 8.3 of 32 lanes active).
//   * the joint loop is ROLLED.  A fully unrolled 20-joint pass is ~100 KB of SASS and ran at a 51 %
//     instruction-cache hit rate; rolled it is a few KB.  Per-lane joint vectors (line-search origin
//     xold, direction p, previous interpolated solution prev, Jacobian columns) live in shared memory
//     in [element][lane] order (bank = lane: conflict-free), indexed by the rolled joint counter; the
//     chain constants are read from the constant bank with a warp-uniform index.
//   * the trial point is never stored: x_j = clamp(xold_j + alam * p_j) is recomputed where needed.
//
// Restates (reference file:line in the reference repository): NewtonRoot::GlobalSolve as driven by
// IKDB/Klamp't (call sites grr/graph.py:779-790, grr/solver.py:799,821,884).
#pragma once
#include "grr_device.cuh"

namespace grr {

using Chain32f = ChainDev<float, kMaxJ>;

// floats (Reals) of shared memory one lane needs
template <int M> __host__ __device__ constexpr int lane_reals(int n) { return n * (3 + M); }

// Shared-memory view of one lane's joint vectors.  Element k of array `arr` of lane `tid` lives at
// base[(off_arr + k) * T + tid].
template <typename Real, int M>
struct LaneMem {
  Real *base;   // already offset by tid
  int T;        // lanes per CTA (stride)
  int n;
  __device__ __forceinline__ Real &xold(int j) const { return base[j * T]; }
  __device__ __forceinline__ Real &p(int j) const { return base[(n + j) * T]; }
  __device__ __forceinline__ Real &prev(int j) const { return base[(2 * n + j) * T]; }
  __device__ __forceinline__ Real &col(int j, int r) const { return base[(3 * n + j * M + r) * T]; }
};

template <typename Real, int M>
struct LaneIK {
  Real fold, slope, alam, alam2, f2, alamin, stpmax, resid;
  int it, evals, phase;
};

// Start an instance: the seed must already be in mem.xold; p is zeroed so that x = clamp(seed).
template <typename Real, int M>
__device__ __forceinline__ void lane_begin(LaneIK<Real, M> &L) {
  L.it = 0; L.evals = 0; L.phase = 0;
  L.alam = 0; L.alam2 = 0; L.f2 = 0; L.fold = 0; L.slope = 0; L.alamin = 0; L.stpmax = 0; L.resid = 0;
}

template <typename Real, typename ChainT>
__device__ __forceinline__ Real lane_x(const ChainT &c, Real xold, Real p, Real alam, int j) {
  return min_(max_(xold + alam * p, c.bmin[j]), c.bmax[j]);
}

// One trip.  Returns IK_RUNNING or the final status.  After a final status the solution is
// x_j = lane_x(xold_j, p_j, alam) (for IK_STALL the caller must use xold).
template <typename Real, int M, typename ChainT>
__device__ __forceinline__ int lane_trip(const ChainT &c, const Target<Real, M> &tg, LaneIK<Real, M> &L,
                                         const LaneMem<Real, M> &mem) {
  const int n = c.n;
  const Real tolf = c.tol, tolx = Real(0.01) * c.tol, ALF = Real(1e-4);
  // ---- backward chain pass at x = clamp(xold + alam p) ----
  Real v0 = c.ee[0], v1 = c.ee[1], v2 = c.ee[2];
  Real G[9];
#pragma unroll
  for (int k = 0; k < 9; k++) G[k] = c.Rtail[k];
  Real S[M * (M + 1) / 2];
#pragma unroll
  for (int k = 0; k < M * (M + 1) / 2; k++) S[k] = 0;
  Real xn2 = 0;
#pragma unroll 1
  for (int j = n - 1; j >= 0; --j) {
    const Real xj = lane_x<Real>(c, mem.xold(j), mem.p(j), L.alam, j);
    xn2 += xj * xj;
    Real s, co;
    sincos_(xj, &s, &co);
    Real Mj[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Mj[k] = c.A[j][k] + s * c.B[j][k] - co * c.C[j][k];
    const Real a0 = c.ax[j][0], a1 = c.ax[j][1], a2 = c.ax[j][2];
    const Real w0 = a1 * v2 - a2 * v1, w1 = a2 * v0 - a0 * v2, w2 = a0 * v1 - a1 * v0;
    Real col[M];
#pragma unroll
    for (int r = 0; r < 3; r++) col[r] = G[r] * w0 + G[3 + r] * w1 + G[6 + r] * w2;      // G^T (axis x v)
    if (M == 6) {
#pragma unroll
      for (int r = 0; r < 3; r++) col[3 + r] = G[r] * a0 + G[3 + r] * a1 + G[6 + r] * a2;  // G^T axis
    }
#pragma unroll
    for (int r = 0; r < M; r++) {
      mem.col(j, r) = col[r];
#pragma unroll
      for (int q = 0; q <= r; q++) S[tri(r, q)] += col[r] * col[q];
    }
    const Real n0 = Mj[0] * v0 + Mj[1] * v1 + Mj[2] * v2 + c.tp[j][0];
    const Real n1 = Mj[3] * v0 + Mj[4] * v1 + Mj[5] * v2 + c.tp[j][1];
    const Real n2 = Mj[6] * v0 + Mj[7] * v1 + Mj[8] * v2 + c.tp[j][2];
    v0 = n0; v1 = n1; v2 = n2;
    Real nG[9];
    m3_mul(Mj, G, nG);
#pragma unroll
    for (int k = 0; k < 9; k++) G[k] = nG[k];
  }
  L.evals++;
  Real Fw[M], FL[M];
  Fw[0] = v0 - tg.x[0]; Fw[1] = v1 - tg.x[1]; Fw[2] = v2 - tg.x[2];
  Real f = Fw[0] * Fw[0] + Fw[1] * Fw[1] + Fw[2] * Fw[2];
  m3_Tvec(G, Fw, FL);
  if (M == 6) {
    Real Re[9], e[3], ew[3];
    m3_Tmul(tg.R, G, Re);
    so3_log(Re, e);
    m3_vec(G, e, ew);
    FL[M - 3] = e[0]; FL[M - 2] = e[1]; FL[M - 1] = e[2];
    Fw[M - 3] = ew[0]; Fw[M - 2] = ew[1]; Fw[M - 1] = ew[2];
    f += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  }
  f *= Real(0.5);
  Real fmax = 0;
#pragma unroll
  for (int r = 0; r < M; r++) fmax = max_(fmax, abs_(Fw[r]));
  L.resid = fmax;

  bool check_convx = false;
  if (L.phase == 0) {
    if (fmax < tolf) return IK_OK;
    if (c.max_iters <= 0) return IK_MAXIT;
    L.stpmax = Real(10) * max_(sqrt_(xn2), Real(n));
  } else {
    if (L.alam < L.alamin) { L.it++; L.alam = 0; return IK_STALL; }   // x <- xold
    if (!(f <= L.fold + ALF * L.alam * L.slope)) {
      Real tmplam;
      if (L.alam == Real(1)) tmplam = -L.slope / (Real(2) * (f - L.fold - L.slope));
      else {
        const Real rhs1 = f - L.fold - L.alam * L.slope, rhs2 = L.f2 - L.fold - L.alam2 * L.slope;
        const Real a = (rhs1 / (L.alam * L.alam) - rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
        const Real b = (-L.alam2 * rhs1 / (L.alam * L.alam) + L.alam * rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
        if (a == Real(0)) tmplam = -L.slope / (Real(2) * b);
        else {
          const Real disc = b * b - Real(3) * a * L.slope;
          if (disc < Real(0)) tmplam = Real(0.5) * L.alam;
          else if (b <= Real(0)) tmplam = (-b + sqrt_(disc)) / (Real(3) * a);
          else tmplam = -L.slope / (b + sqrt_(disc));
        }
        if (tmplam > Real(0.5) * L.alam) tmplam = Real(0.5) * L.alam;
      }
      L.alam2 = L.alam; L.f2 = f;
      L.alam = max_(tmplam, Real(0.1) * L.alam);
      return IK_RUNNING;
    }
    L.it++;
    if (fmax < tolf) return IK_OK;
    check_convx = true;
  }

  // ---- new Newton step from the columns of this pass ----
  Real y[M];
  {
    Real A[M * (M + 1) / 2];
    const Real l2 = c.lambda * c.lambda;
    if (M == 3) {
#pragma unroll
      for (int k = 0; k < 6; k++) A[k] = S[k];
#pragma unroll
      for (int r = 0; r < M; r++) A[tri(r, r)] += l2;
      chol_solve<Real, M>(A, FL, y);
    } else {
      Real D[9], Src[9], Srr[9], DSrc[9], DSrr[9];
      so3_Jl_inv(&FL[M - 3], D);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) {
          Src[3 * r + q] = S[tri(3 + r, q)];
          Srr[3 * r + q] = (r >= q) ? S[tri(3 + r, 3 + q)] : S[tri(3 + q, 3 + r)];
        }
      m3_mul(D, Src, DSrc);
      m3_mul(D, Srr, DSrr);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q <= r; q++) A[tri(r, q)] = S[tri(r, q)];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) A[tri(3 + r, q)] = DSrc[3 * r + q];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q <= r; q++)
          A[tri(3 + r, 3 + q)] = DSrr[3 * r] * D[3 * q] + DSrr[3 * r + 1] * D[3 * q + 1] + DSrr[3 * r + 2] * D[3 * q + 2];
#pragma unroll
      for (int r = 0; r < M; r++) A[tri(r, r)] += l2;
      chol_solve<Real, M>(A, FL, y);
      Real z[3];
      m3_Tvec(D, &y[M - 3], z);
      y[M - 3] = z[0]; y[M - 2] = z[1]; y[M - 1] = z[2];
    }
  }
  Real pn = 0, slope = 0, test = 0, testx = 0;
#pragma unroll 1
  for (int j = 0; j < n; j++) {
    const Real xo = mem.xold(j);
    const Real xj = lane_x<Real>(c, xo, mem.p(j), L.alam, j);
    testx = max_(testx, abs_(xj - xo) / max_(abs_(xj), Real(1)));
    Real pj = 0, gj = 0;
#pragma unroll
    for (int r = 0; r < M; r++) { const Real cj = mem.col(j, r); pj += cj * y[r]; gj += cj * FL[r]; }
    pj = -pj;
    mem.xold(j) = xj;
    mem.p(j) = pj;
    pn += pj * pj;
    slope += gj * pj;
    test = max_(test, abs_(pj) / max_(abs_(xj), Real(1)));
  }
  L.alam = 0;                                     // x == xold (the accepted point) until the next trial is armed
  if (check_convx) {
    if (testx < tolx) return IK_CONVX;
    if (L.it >= c.max_iters) return IK_MAXIT;
  }
  pn = sqrt_(pn);
  if (pn > L.stpmax) {
    const Real sc = L.stpmax / pn;
    test = 0;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      const Real pj = mem.p(j) * sc;
      mem.p(j) = pj;
      test = max_(test, abs_(pj) / max_(abs_(mem.xold(j)), Real(1)));
    }
    slope *= sc;
  }
  if (!(slope < Real(0))) { L.it++; return IK_STALL; }
  L.alamin = tolx / test;
  L.fold = f; L.slope = slope; L.alam = 1; L.alam2 = 0; L.f2 = 0;
  L.phase = 1;
  return IK_RUNNING;
}

}  // namespace grr
