// grr_device.cuh -- device-side building blocks of the B200 roadmap-construction path.
//
// Everything is templated on the arithmetic type (float = product path on the FP32 SIMT pipe,
// double = verification build) and on the residual rows M (3 = free orientation, 6 = fixed /
// variable).  The joint count is a run-time value (<= 32): the joint loop is rolled on purpose,
// see grr_lane.cuh.
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
// iteration itself (a resumable per-lane state machine) is in grr_lane.cuh.
#pragma once
#include <stdio.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

// Index checks of our own for the kernels with data-dependent addressing (the global pair queue, the CSP
// sweeps): compiled in with -DGRR_DEBUG_CHECKS (`make debug` -> libgrr_b200_dbg.so, loaded with GRR_LIB=...);
// a failed check traps, which surfaces as a launch error of the ABI call.
#ifdef GRR_DEBUG_CHECKS
#define GRR_DCHECK(cond) do { if (!(cond)) { printf("GRR_DCHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define GRR_DCHECK(cond) do { } while (0)
#endif

// Lane-machine building blocks are __host__ __device__: the product runs them in the kernels only; the host
// instantiation exists for the near-tie study / CPU checks of the same source (tools/lane_host.cu, tests).
#define GRR_DEV __host__ __device__ __forceinline__

namespace grr {

constexpr int kMaxJ = 32;

// ------------------------------------------------------------------------------------------
// scalar helpers, overloaded on precision
// ------------------------------------------------------------------------------------------
GRR_DEV void sincos_(float a, float *s, float *c) { sincosf(a, s, c); }
GRR_DEV void sincos_(double a, double *s, double *c) { sincos(a, s, c); }
GRR_DEV float sqrt_(float a) { return sqrtf(a); }
GRR_DEV double sqrt_(double a) { return sqrt(a); }
GRR_DEV float atan2_(float a, float b) { return atan2f(a, b); }
GRR_DEV double atan2_(double a, double b) { return atan2(a, b); }
GRR_DEV float abs_(float a) { return fabsf(a); }
GRR_DEV double abs_(double a) { return fabs(a); }
GRR_DEV float max_(float a, float b) { return fmaxf(a, b); }
GRR_DEV double max_(double a, double b) { return fmax(a, b); }
GRR_DEV float fma_(float a, float b, float c) { return fmaf(a, b, c); }
GRR_DEV double fma_(double a, double b, double c) { return fma(a, b, c); }
GRR_DEV float min_(float a, float b) { return fminf(a, b); }
GRR_DEV double min_(double a, double b) { return fmin(a, b); }
GRR_DEV float ceil_(float a) { return ceilf(a); }
GRR_DEV double ceil_(double a) { return ceil(a); }
GRR_DEV float fmod_(float a, float b) { return fmodf(a, b); }
GRR_DEV double fmod_(double a, double b) { return fmod(a, b); }

// ---- 'spin' joints (continuous rotation; Klamp't joint type 'spin', reference flag graph.py:556-561) ----------------
// robot.distance takes the wrapped angular difference of such a joint and robot.interpolate its shorter arc (SURVEY K1/K2,
// A.1: [recalled] KrisLibrary AngleNormalize / AngleDiff / AngleInterp); the oracle restates the same three functions.
// spin_mask bit j of the chain constants marks active joint j; every use is behind `if (spin_mask ...)`, so chains of
// 'normal' joints (every shipped problem) run the code they ran before.
template <typename Real> GRR_DEV Real angle_normalize(Real a) {           // into [0, 2 pi)
  const Real an = fmod_(a, Real(6.283185307179586476925286766559));
  return an < Real(0) ? an + Real(6.283185307179586476925286766559) : an;
}
template <typename Real> GRR_DEV Real angle_diff(Real a, Real b) {        // a - b, into (-pi, pi]
  const Real d = angle_normalize(angle_normalize(a) - angle_normalize(b));
  return d > Real(3.1415926535897932384626433832795) ? d - Real(6.283185307179586476925286766559) : d;
}
// value of active joint j as an IK solve returns it ([recalled] robot.NormalizeAngles after the solve)
template <typename Real> GRR_DEV Real joint_out(unsigned spin_mask, int j, Real v) {
  return (spin_mask >> j & 1u) ? angle_normalize(v) : v;
}
// joint-space difference a - b of active joint j as robot.distance measures it
template <typename Real> GRR_DEV Real joint_diff(unsigned spin_mask, int j, Real a, Real b) {
  return (spin_mask >> j & 1u) ? angle_diff(a, b) : a - b;
}

// ------------------------------------------------------------------------------------------
// chain constants (kernel parameter, __grid_constant__: lives in the constant bank)
// ------------------------------------------------------------------------------------------
// per-joint constants every joint kind needs, packed for two 16-byte constant loads
template <typename Real>
struct alignas(16) JointConst {
  Real tp[3];      // translation preceding the joint
  Real sgn;        // +-1 for the short kinds (axis = sgn e_K)
  Real bmin, bmax; // joint box (+-inf for revolute joints with range >= 2 pi)
  int kind;        // 0 generic; 1/2/3: identity preceding rotation and axis = +-e_x / e_y / e_z
  int spin;        // 1: 'spin' joint (unbounded; wrapped distance, shorter-arc interpolation)
};

template <typename Real, int NJ>
struct ChainDev {
  int n, mode, nlinks_eps, max_iters;
  int small_angles;            // every joint limit inside [-pi/4, pi/4]: sincos needs no range reduction
  int axes_positive;           // every short-path joint has sgn = +1
  unsigned spin_mask;          // bit j: active joint j is a 'spin' joint (0 for every shipped problem)
  Real tol, lambda;
  // near-tie margins of the single-precision edge phase (FLAG builds of lane_step / pair_advance; all 0 otherwise)
  Real fl_f, fl_f2, fl_x, fl_alam, fl_rel, fl_leaf, fl_nd, fl_dd, fl_curv, fl_curv1, fl_jac;
  int fl_maxit;                // solves longer than this many Newton iterations are not trusted in single precision
  // M_j(theta) = Rp_j * Rot(axis_j, theta) = A + sin(theta) * B - cos(theta) * C
  Real A[NJ][9], B[NJ][9], C[NJ][9];
  Real tp[NJ][3];
  Real ax[NJ][3];
  Real bmin[NJ], bmax[NJ];     // +-inf for revolute joints with range >= 2 pi
  JointConst<Real> jc[NJ];
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
GRR_DEV void m3_mul(const Real *A, const Real *B, Real *C) {  // C = A B (no alias)
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename Real>
GRR_DEV void m3_Tmul(const Real *A, const Real *B, Real *C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
template <typename Real>
GRR_DEV void m3_vec(const Real *A, const Real *v, Real *r) {  // r = A v (no alias)
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
template <typename Real>
GRR_DEV void m3_Tvec(const Real *A, const Real *v, Real *r) {  // r = A^T v
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}

// klampt.math.so3.from_moment: exp map (graph.py:96,152)
template <typename Real>
__host__ __device__ void so3_exp(const Real *w, Real *R) {
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
__host__ __device__ void so3_log(const Real *R, Real *w) {
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
__host__ __device__ void so3_Jl_inv(const Real *e, Real *D) {
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
GRR_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t *out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
    uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }  // packed lower triangle, r >= c

// forward-facing FK for the API (EE point + goal-link rotation), same backward walk
template <typename Real, int NJ>
__host__ __device__ void chain_fk(const ChainDev<Real, NJ> &c, const Real *x, Real *p, Real *Rl) {
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
GRR_DEV void chol_solve(Real (&A)[M * (M + 1) / 2], const Real *b, Real *y) {
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

enum : int { IK_RUNNING = -1, IK_OK = 0, IK_STALL = 1, IK_CONVX = 2, IK_MAXIT = 3 };

// ------------------------------------------------------------------------------------------
// workspace targets: setIKProblem (graph.py:737-745) and workspaceInterpolate (graph.py:768-777)
// ------------------------------------------------------------------------------------------
template <typename Real, int M, int NJ>
GRR_DEV void target_from_params(const ChainDev<Real, NJ> &c, const Real *w, Target<Real, M> &tg) {
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
GRR_DEV void winterp_begin(const ChainDev<Real, NJ> &c, const Real *wa, const Real *wb, WInterp<Real, M> &W) {
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
GRR_DEV void winterp_at(const WInterp<Real, M> &W, Real u, Target<Real, M> &tg) {
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
