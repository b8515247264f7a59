// grr_device.cuh -- device-side building blocks of the B200 roadmap-construction path.
//
// Everything is templated on the arithmetic type (float = product path on the FP32 SIMT pipe,
// double = verification build), on the number of active joints N (compile-time so that the
// per-thread joint arrays live in registers and every chain constant is a constant-bank
// operand; N = 0 selects the runtime-n fallback), and on the residual rows M (3 = free
// orientation, 6 = fixed / variable).
//
// What is restated here (reference file:line in the reference repository):
//   * IK residual / Jacobian of a serial chain  (Klamp't RobotIKFunction; call sites
//     grr/graph.py:779-790, grr/solver.py:799,821,884)
//   * NewtonRoot::GlobalSolve damped Newton + NR line search with box clamp
//   * validEdgeLinear                                  (grr/graph.py:938-1006)
//   * IKTask.setParams / interpolate                   (grr/graph.py:84-103,143-164)
//
// Formulation (B200-first, not a translation): one thread owns one IK instance and walks the
// chain BACKWARDS (goal link -> base).  The backward walk yields the end-effector point, the
// goal-link rotation and every Jacobian column expressed in the goal-link frame in a single
// pass with O(1) live state (v, G), so no per-joint frames are stored; the damped step
// p = -J^T (J J^T + lambda^2 I)^-1 F is invariant under that change of frame.  The Newton
// iteration is written as a resumable state machine (one chain pass per "trip") so that a
// lane that finishes an instance can start the next one without waiting for its warp.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace grr {

constexpr int kMaxJ = 32;

// ------------------------------------------------------------------------------------------
// scalar helpers, overloaded on precision
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_(float a, float *s, float *c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ float sqrt_(float a) { return sqrtf(a); }
__device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
__device__ __forceinline__ float atan2_(float a, float b) { return atan2f(a, b); }
__device__ __forceinline__ double atan2_(double a, double b) { return atan2(a, b); }
__device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_(double a) { return fabs(a); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float ceil_(float a) { return ceilf(a); }
__device__ __forceinline__ double ceil_(double a) { return ceil(a); }
template <typename Real> __device__ __forceinline__ Real inf_();
template <> __device__ __forceinline__ float inf_<float>() { return CUDART_INF_F; }
template <> __device__ __forceinline__ double inf_<double>() { return CUDART_INF; }

// ------------------------------------------------------------------------------------------
// chain constants (kernel parameter, __grid_constant__: lives in the constant bank)
// ------------------------------------------------------------------------------------------
template <typename Real, int NJ>
struct ChainDev {
  int n, mode, nlinks_eps, max_iters;
  Real tol, lambda;
  // M_j(theta) = Rp_j * Rot(axis_j, theta) = A + sin(theta) * B - cos(theta) * C
  Real A[NJ][9], B[NJ][9], C[NJ][9];
  Real tp[NJ][3];
  Real ax[NJ][3];
  Real bmin[NJ], bmax[NJ];     // +-inf for revolute joints with range >= 2 pi
  float slo[NJ], srange[NJ];   // random-start sampling box, single precision by specification
  Real ee[3];
  Real Rtail[9];
  Real Rgoal[9];
};

template <typename Real, int M>
struct Target {
  Real x[3];
  Real R[M == 6 ? 9 : 1];
};

// ------------------------------------------------------------------------------------------
// 3x3 helpers (row-major)
// ------------------------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ void m3_mul(const Real *A, const Real *B, Real *C) {  // C = A B (no alias)
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename Real>
__device__ __forceinline__ void m3_Tmul(const Real *A, const Real *B, Real *C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
template <typename Real>
__device__ __forceinline__ void m3_vec(const Real *A, const Real *v, Real *r) {  // r = A v (no alias)
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
template <typename Real>
__device__ __forceinline__ void m3_Tvec(const Real *A, const Real *v, Real *r) {  // r = A^T v
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}

// klampt.math.so3.from_moment: exp map (graph.py:96,152)
template <typename Real>
__device__ void so3_exp(const Real *w, Real *R) {
  Real t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  Real th = sqrt_(t2);
  if (th < Real(1e-12)) {
    R[0] = 1; R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2]; R[4] = 1; R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0]; R[8] = 1;
    return;
  }
  Real a0 = w[0] / th, a1 = w[1] / th, a2 = w[2] / th, s, c;
  sincos_(th, &s, &c);
  Real v = Real(1) - c;
  R[0] = c + v * a0 * a0;      R[1] = v * a0 * a1 - s * a2; R[2] = v * a0 * a2 + s * a1;
  R[3] = v * a1 * a0 + s * a2; R[4] = c + v * a1 * a1;      R[5] = v * a1 * a2 - s * a0;
  R[6] = v * a2 * a0 - s * a1; R[7] = v * a2 * a1 + s * a0; R[8] = c + v * a2 * a2;
}

// klampt.math.so3.moment: log map with theta = atan2(|skew|, (tr-1)/2) (graph.py:118,154)
template <typename Real>
__device__ void so3_log(const Real *R, Real *w) {
  Real v0 = Real(0.5) * (R[7] - R[5]), v1 = Real(0.5) * (R[2] - R[6]), v2 = Real(0.5) * (R[3] - R[1]);
  Real s = sqrt_(v0 * v0 + v1 * v1 + v2 * v2);
  Real c = Real(0.5) * (R[0] + R[4] + R[8] - Real(1));
  Real th = atan2_(s, c);
  if (c > Real(-0.9)) {
    Real k = (s > Real(1e-8)) ? th / s : Real(1);
    w[0] = v0 * k; w[1] = v1 * k; w[2] = v2 * k;
    return;
  }
  Real omc = Real(1) - c;
  Real x = sqrt_(max_((R[0] - c) / omc, Real(0)));
  Real y = sqrt_(max_((R[4] - c) / omc, Real(0)));
  Real z = sqrt_(max_((R[8] - c) / omc, Real(0)));
  if (x >= y && x >= z) {
    if (v0 < 0) x = -x;
    if ((R[1] + R[3]) * x < 0) y = -y;
    if ((R[2] + R[6]) * x < 0) z = -z;
  } else if (y >= z) {
    if (v1 < 0) y = -y;
    if ((R[1] + R[3]) * y < 0) x = -x;
    if ((R[5] + R[7]) * y < 0) z = -z;
  } else {
    if (v2 < 0) z = -z;
    if ((R[2] + R[6]) * z < 0) x = -x;
    if ((R[5] + R[7]) * z < 0) y = -y;
  }
  Real nn = sqrt_(x * x + y * y + z * z);
  if (nn < Real(1e-30)) nn = 1;
  Real k = th / nn;
  w[0] = x * k; w[1] = y * k; w[2] = z * k;
}

// inverse left Jacobian of SO(3) at e: D = I - 0.5 [e]x + g [e]x^2
template <typename Real>
__device__ void so3_Jl_inv(const Real *e, Real *D) {
  Real t2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  Real g;
  if (t2 < Real(1e-8)) g = Real(1.0 / 12.0);
  else {
    Real h = Real(0.5) * sqrt_(t2), sh, ch;
    sincos_(h, &sh, &ch);
    g = (Real(1) - h * ch / sh) / t2;
  }
  D[0] = 1 + g * (e[0] * e[0] - t2);             D[1] = Real(0.5) * e[2] + g * e[0] * e[1];  D[2] = Real(-0.5) * e[1] + g * e[0] * e[2];
  D[3] = Real(-0.5) * e[2] + g * e[1] * e[0];    D[4] = 1 + g * (e[1] * e[1] - t2);          D[5] = Real(0.5) * e[0] + g * e[1] * e[2];
  D[6] = Real(0.5) * e[1] + g * e[2] * e[0];     D[7] = Real(-0.5) * e[0] + g * e[2] * e[1]; D[8] = 1 + g * (e[2] * e[2] - t2);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG shared by specification with the CPU oracle)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t *out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ------------------------------------------------------------------------------------------
// chain pass: residual, merit and (optionally) Jacobian columns + normal matrix, one backward walk
// ------------------------------------------------------------------------------------------
template <int N> struct NA_ { static constexpr int v = (N > 0 ? N : kMaxJ); };

template <typename Real, int N, int M>
struct PassOut {
  Real Fw[M];            // residual, world frame (stopping test is max|Fw|)
  Real FL[M];            // residual, goal-link frame (used by the solve)
  Real f;                // 0.5 |F|^2
  Real cp[NA_<N>::v][3]; // position Jacobian columns, goal-link frame
  Real cr[M == 6 ? NA_<N>::v : 1][3];  // raw rotation columns (G_j^T axis_j), goal-link frame
  Real S[M * (M + 1) / 2];             // packed lower triangle of [cp;cr][cp;cr]^T summed over joints
};

__host__ __device__ constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }  // r >= c

template <typename Real, int N, int M, int NJ, bool WITH_J>
__device__ __forceinline__ void chain_pass(const ChainDev<Real, NJ> &c, const Real *x, const Target<Real, M> &tg,
                                           PassOut<Real, N, M> &o) {
  const int n = (N > 0 ? N : c.n);
  Real v[3] = {c.ee[0], c.ee[1], c.ee[2]};
  Real G[9];
  constexpr bool needG = WITH_J || (M == 6);
  if (needG) {
#pragma unroll
    for (int k = 0; k < 9; k++) G[k] = c.Rtail[k];
  }
  if (WITH_J) {
#pragma unroll
    for (int k = 0; k < M * (M + 1) / 2; k++) o.S[k] = 0;
  }
#pragma unroll
  for (int jj = 0; jj < NA_<N>::v; jj++) {
    const int j = n - 1 - jj;
    if (N == 0 && j < 0) break;
    Real s, co;
    sincos_(x[j], &s, &co);
    Real Mj[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Mj[k] = c.A[j][k] + s * c.B[j][k] - co * c.C[j][k];
    if (WITH_J) {
      Real w[3] = {c.ax[j][1] * v[2] - c.ax[j][2] * v[1], c.ax[j][2] * v[0] - c.ax[j][0] * v[2],
                   c.ax[j][0] * v[1] - c.ax[j][1] * v[0]};
      Real col[6];
      m3_Tvec(G, w, col);
      o.cp[j][0] = col[0]; o.cp[j][1] = col[1]; o.cp[j][2] = col[2];
      if (M == 6) {
        Real a[3] = {c.ax[j][0], c.ax[j][1], c.ax[j][2]};
        m3_Tvec(G, a, col + 3);
        o.cr[j][0] = col[3]; o.cr[j][1] = col[4]; o.cr[j][2] = col[5];
      }
#pragma unroll
      for (int r = 0; r < M; r++)
#pragma unroll
        for (int q = 0; q <= r; q++) o.S[tri(r, q)] += col[r] * col[q];
    }
    Real nv[3];
    m3_vec(Mj, v, nv);
    v[0] = nv[0] + c.tp[j][0]; v[1] = nv[1] + c.tp[j][1]; v[2] = nv[2] + c.tp[j][2];
    if (needG) {
      Real nG[9];
      m3_mul(Mj, G, nG);
#pragma unroll
      for (int k = 0; k < 9; k++) G[k] = nG[k];
    }
  }
  o.Fw[0] = v[0] - tg.x[0]; o.Fw[1] = v[1] - tg.x[1]; o.Fw[2] = v[2] - tg.x[2];
  Real f = o.Fw[0] * o.Fw[0] + o.Fw[1] * o.Fw[1] + o.Fw[2] * o.Fw[2];
  if (needG) m3_Tvec(G, o.Fw, o.FL);
  if (M == 6) {
    Real Re[9], e[3], ew[3];
    m3_Tmul(tg.R, G, Re);          // R_goal^T R_link
    so3_log(Re, e);
    m3_vec(G, e, ew);              // world-frame rotation error = log(R_link R_goal^T)
    o.FL[M - 3] = e[0]; o.FL[M - 2] = e[1]; o.FL[M - 1] = e[2];
    o.Fw[M - 3] = ew[0]; o.Fw[M - 2] = ew[1]; o.Fw[M - 1] = ew[2];
    f += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  }
  o.f = Real(0.5) * f;
}

// forward-facing FK for the API (EE point + goal-link rotation), same backward walk
template <typename Real, int NJ>
__device__ void chain_fk(const ChainDev<Real, NJ> &c, const Real *x, Real *p, Real *Rl) {
  Real v[3] = {c.ee[0], c.ee[1], c.ee[2]}, G[9];
  for (int k = 0; k < 9; k++) G[k] = c.Rtail[k];
  for (int j = c.n - 1; j >= 0; j--) {
    Real s, co, Mj[9], nv[3], nG[9];
    sincos_(x[j], &s, &co);
    for (int k = 0; k < 9; k++) Mj[k] = c.A[j][k] + s * c.B[j][k] - co * c.C[j][k];
    m3_vec(Mj, v, nv);
    v[0] = nv[0] + c.tp[j][0]; v[1] = nv[1] + c.tp[j][1]; v[2] = nv[2] + c.tp[j][2];
    m3_mul(Mj, G, nG);
    for (int k = 0; k < 9; k++) G[k] = nG[k];
  }
  p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
  if (Rl) for (int k = 0; k < 9; k++) Rl[k] = G[k];
}

// ------------------------------------------------------------------------------------------
// damped step: y = (J J^T + lambda^2 I)^-1 F_L  via Cholesky; p_j = -(cp_j . y_p + cr_j . z)
// ------------------------------------------------------------------------------------------
template <typename Real, int M>
__device__ __forceinline__ void chol_solve(Real (&A)[M * (M + 1) / 2], const Real *b, Real *y) {
#pragma unroll
  for (int r = 0; r < M; r++) {
#pragma unroll
    for (int s = 0; s <= r; s++) {
      Real acc = A[tri(r, s)];
#pragma unroll
      for (int k = 0; k < s; k++) acc -= A[tri(r, k)] * A[tri(s, k)];
      if (r == s) A[tri(r, r)] = sqrt_(acc);
      else A[tri(r, s)] = acc / A[tri(s, s)];
    }
  }
#pragma unroll
  for (int r = 0; r < M; r++) {
    Real acc = b[r];
#pragma unroll
    for (int k = 0; k < r; k++) acc -= A[tri(r, k)] * y[k];
    y[r] = acc / A[tri(r, r)];
  }
#pragma unroll
  for (int r = M - 1; r >= 0; r--) {
    Real acc = y[r];
#pragma unroll
    for (int k = r + 1; k < M; k++) acc -= A[tri(k, r)] * y[k];
    y[r] = acc / A[tri(r, r)];
  }
}

// ------------------------------------------------------------------------------------------
// IK lane: resumable NewtonRoot::GlobalSolve
// ------------------------------------------------------------------------------------------
enum : int { IK_RUNNING = -1, IK_OK = 0, IK_STALL = 1, IK_CONVX = 2, IK_MAXIT = 3 };

template <typename Real, int N, int M>
struct IkLane {
  static constexpr int NA = NA_<N>::v;
  Real x[NA];       // current (trial) point
  Real xold[NA];    // start of the current line search
  Real p[NA];       // current Newton direction
  Real fold, slope, alam, alam2, f2, alamin, stpmax, resid;
  int it;           // completed Newton iterations
  int evals;        // chain passes
  int phase;        // 0 = first evaluation at the seed, 1 = line-search trial
};

// Clamp the seed into the joint box and reset the state machine (IKDB solve, startRandom=False
// path: robot.setConfig(seed) then solver.solve(); graph.py:783-790).
template <typename Real, int N, int M, int NJ>
__device__ __forceinline__ void ik_begin(const ChainDev<Real, NJ> &c, IkLane<Real, N, M> &L) {
  const int n = (N > 0 ? N : c.n);
#pragma unroll
  for (int j = 0; j < IkLane<Real, N, M>::NA; j++) {
    if (N == 0 && j >= n) break;
    L.x[j] = min_(max_(L.x[j], c.bmin[j]), c.bmax[j]);
  }
  L.it = 0; L.evals = 0; L.phase = 0;
  L.alam = 1; L.alam2 = 0; L.f2 = 0; L.fold = 0; L.slope = 0; L.alamin = 0; L.stpmax = 0; L.resid = 0;
}

// One trip = one chain pass at L.x + the state-machine transition.  Returns IK_RUNNING or a final
// status; on return with a final status L.x holds the result configuration.
template <typename Real, int N, int M, int NJ>
__device__ __forceinline__ int ik_trip(const ChainDev<Real, NJ> &c, const Target<Real, M> &tg, IkLane<Real, N, M> &L) {
  constexpr int NA = IkLane<Real, N, M>::NA;
  const int n = (N > 0 ? N : c.n);
  const Real tolf = c.tol, tolx = Real(0.01) * c.tol, ALF = Real(1e-4);
  PassOut<Real, N, M> o;
  chain_pass<Real, N, M, NJ, true>(c, L.x, tg, o);
  L.evals++;
  Real fmax = 0;
#pragma unroll
  for (int r = 0; r < M; r++) fmax = max_(fmax, abs_(o.Fw[r]));
  L.resid = fmax;

  if (L.phase == 0) {
    if (fmax < tolf) return IK_OK;
    if (c.max_iters <= 0) return IK_MAXIT;
    Real xn = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; xn += L.x[j] * L.x[j]; }
    L.stpmax = Real(10) * max_(sqrt_(xn), Real(n));
  } else {
    if (L.alam < L.alamin) {           // line search stalled: x <- xold, failure
#pragma unroll
      for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; L.x[j] = L.xold[j]; }
      L.it++;
      return IK_STALL;
    }
    if (!(o.f <= L.fold + ALF * L.alam * L.slope)) {
      // backtrack (Numerical Recipes lnsrch)
      Real tmplam;
      if (L.alam == Real(1)) tmplam = -L.slope / (Real(2) * (o.f - L.fold - L.slope));
      else {
        Real rhs1 = o.f - L.fold - L.alam * L.slope, rhs2 = L.f2 - L.fold - L.alam2 * L.slope;
        Real a = (rhs1 / (L.alam * L.alam) - rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
        Real b = (-L.alam2 * rhs1 / (L.alam * L.alam) + L.alam * rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
        if (a == Real(0)) tmplam = -L.slope / (Real(2) * b);
        else {
          Real disc = b * b - Real(3) * a * L.slope;
          if (disc < Real(0)) tmplam = Real(0.5) * L.alam;
          else if (b <= Real(0)) tmplam = (-b + sqrt_(disc)) / (Real(3) * a);
          else tmplam = -L.slope / (b + sqrt_(disc));
        }
        if (tmplam > Real(0.5) * L.alam) tmplam = Real(0.5) * L.alam;
      }
      L.alam2 = L.alam; L.f2 = o.f;
      L.alam = max_(tmplam, Real(0.1) * L.alam);
#pragma unroll
      for (int j = 0; j < NA; j++) {
        if (N == 0 && j >= n) break;
        L.x[j] = min_(max_(L.xold[j] + L.alam * L.p[j], c.bmin[j]), c.bmax[j]);
      }
      return IK_RUNNING;
    }
    // trial accepted: iteration complete
    L.it++;
    if (fmax < tolf) return IK_OK;
    Real test = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) {
      if (N == 0 && j >= n) break;
      test = max_(test, abs_(L.x[j] - L.xold[j]) / max_(abs_(L.x[j]), Real(1)));
    }
    if (test < tolx) return IK_CONVX;
    if (L.it >= c.max_iters) return IK_MAXIT;
  }

  // ---- new Newton step at x (Jacobian columns of this pass) ----
  Real y[M];
  {
    Real A[M * (M + 1) / 2];
    const Real l2 = c.lambda * c.lambda;
    if (M == 3) {
#pragma unroll
      for (int k = 0; k < 6; k++) A[k] = o.S[k];
    } else {
      // rows 3..5 of J are D * cr with D = Jl^-1(e): transform the raw normal matrix
      Real D[9];
      so3_Jl_inv(&o.FL[M - 3], D);
      // Scc stays; Scr' = Scr D^T (stored as lower block rows 3..5, cols 0..2: (D Src)); Srr' = D Srr D^T
      Real Src[9], Srr[9];   // Src[r][q] = S[3+r][q], full Srr
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) {
          Src[3 * r + q] = o.S[tri(3 + r, q)];
          Srr[3 * r + q] = (r >= q) ? o.S[tri(3 + r, 3 + q)] : o.S[tri(3 + q, 3 + r)];
        }
      Real DSrc[9], DSrr[9];
      m3_mul(D, Src, DSrc);
      m3_mul(D, Srr, DSrr);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q <= r; q++) A[tri(r, q)] = o.S[tri(r, q)];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) A[tri(3 + r, q)] = DSrc[3 * r + q];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q <= r; q++)
          A[tri(3 + r, 3 + q)] = DSrr[3 * r] * D[3 * q] + DSrr[3 * r + 1] * D[3 * q + 1] + DSrr[3 * r + 2] * D[3 * q + 2];
      // z = D^T y_r is applied after the solve
      (void)Src;
    }
#pragma unroll
    for (int r = 0; r < M; r++) A[tri(r, r)] += l2;
    chol_solve<Real, M>(A, o.FL, y);
    if (M == 6) {
      Real D[9], z[3];
      so3_Jl_inv(&o.FL[M - 3], D);
      m3_Tvec(D, &y[M - 3], z);
      y[M - 3] = z[0]; y[M - 2] = z[1]; y[M - 1] = z[2];
    }
  }
  Real pn = 0, slope = 0, test = 0;
#pragma unroll
  for (int j = 0; j < NA; j++) {
    if (N == 0 && j >= n) break;
    Real pj = o.cp[j][0] * y[0] + o.cp[j][1] * y[1] + o.cp[j][2] * y[2];
    Real gj = o.cp[j][0] * o.FL[0] + o.cp[j][1] * o.FL[1] + o.cp[j][2] * o.FL[2];
    if (M == 6) {
      pj += o.cr[j][0] * y[M - 3] + o.cr[j][1] * y[M - 2] + o.cr[j][2] * y[M - 1];
      gj += o.cr[j][0] * o.FL[M - 3] + o.cr[j][1] * o.FL[M - 2] + o.cr[j][2] * o.FL[M - 1];
    }
    pj = -pj;
    L.p[j] = pj;
    pn += pj * pj;
    slope += gj * pj;
  }
  pn = sqrt_(pn);
  if (pn > L.stpmax) {
    const Real sc = L.stpmax / pn;
#pragma unroll
    for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; L.p[j] *= sc; }
    slope *= sc;
  }
  if (!(slope < Real(0))) { L.it++; return IK_STALL; }
#pragma unroll
  for (int j = 0; j < NA; j++) {
    if (N == 0 && j >= n) break;
    test = max_(test, abs_(L.p[j]) / max_(abs_(L.x[j]), Real(1)));
  }
  L.alamin = tolx / test;
  L.fold = o.f; L.slope = slope; L.alam = 1; L.alam2 = 0; L.f2 = 0;
#pragma unroll
  for (int j = 0; j < NA; j++) {
    if (N == 0 && j >= n) break;
    L.xold[j] = L.x[j];
    L.x[j] = min_(max_(L.x[j] + L.p[j], c.bmin[j]), c.bmax[j]);
  }
  L.phase = 1;
  return IK_RUNNING;
}

// Run one instance to completion (seed in L.x).
template <typename Real, int N, int M, int NJ>
__device__ __forceinline__ int ik_solve(const ChainDev<Real, NJ> &c, const Target<Real, M> &tg, IkLane<Real, N, M> &L) {
  ik_begin<Real, N, M, NJ>(c, L);
  int st;
  do { st = ik_trip<Real, N, M, NJ>(c, tg, L); } while (st == IK_RUNNING);
  return st;
}

// ------------------------------------------------------------------------------------------
// workspace targets: setIKProblem (graph.py:737-745) and workspaceInterpolate (graph.py:768-777)
// ------------------------------------------------------------------------------------------
template <typename Real, int M, int NJ>
__device__ __forceinline__ void target_from_params(const ChainDev<Real, NJ> &c, const Real *w, Target<Real, M> &tg) {
  tg.x[0] = w[0]; tg.x[1] = w[1]; tg.x[2] = w[2];
  if (M == 6) {
    if (c.mode == 2) so3_exp(w + 3, tg.R);
    else {
#pragma unroll
      for (int k = 0; k < 9; k++) tg.R[k] = c.Rgoal[k];
    }
  }
}

// Interpolation context of one pair of workspace nodes: position lerp + SO(3) geodesic.
template <typename Real, int M>
struct WInterp {
  Real xa[3], dx[3];
  Real Ra[M == 6 ? 9 : 1];
  Real rel[M == 6 ? 3 : 1];   // log(Ra^T Rb)
  bool rot;                   // variable orientation
};
template <typename Real, int M, int NJ>
__device__ __forceinline__ void winterp_begin(const ChainDev<Real, NJ> &c, const Real *wa, const Real *wb, WInterp<Real, M> &W) {
#pragma unroll
  for (int k = 0; k < 3; k++) { W.xa[k] = wa[k]; W.dx[k] = wb[k] - wa[k]; }
  W.rot = false;
  if (M == 6) {
    if (c.mode == 2) {
      Real Rb[9], Rrel[9];
      so3_exp(wa + 3, W.Ra); so3_exp(wb + 3, Rb);
      m3_Tmul(W.Ra, Rb, Rrel);
      so3_log(Rrel, W.rel);
      W.rot = true;
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) W.Ra[k] = c.Rgoal[k];
    }
  }
}
template <typename Real, int M>
__device__ __forceinline__ void winterp_at(const WInterp<Real, M> &W, Real u, Target<Real, M> &tg) {
#pragma unroll
  for (int k = 0; k < 3; k++) tg.x[k] = W.xa[k] + u * W.dx[k];
  if (M == 6) {
    if (W.rot) {
      Real m[3] = {u * W.rel[0], u * W.rel[1], u * W.rel[2]}, Ru[9];
      so3_exp(m, Ru);
      m3_mul(W.Ra, Ru, tg.R);
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) tg.R[k] = W.Ra[k];
    }
  }
}

}  // namespace grr
