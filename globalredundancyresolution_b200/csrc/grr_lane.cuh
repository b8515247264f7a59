// grr_lane.cuh -- the per-lane IK machine used by every kernel of the path.
//
// Design (B200-first; measured motivation in profiles/r01_edge_kernel.md):
//   * one lane = one IK instance at a time, advanced one "trip" (= one backward chain pass + the
//     Newton / line-search transition) per loop iteration.  Lanes of a warp are never in lock step on
//     an instance: a lane that finishes fetches its next instance immediately, so the only divergent
//     code is the short transition logic, never the chain pass (v1: 8.3 of 32 lanes active).
//   * the joint loop is ROLLED and executed by every lane of the warp unconditionally, so the joint
//     counter is warp-uniform and the chain constants come straight from the constant bank through
//     the uniform datapath.  (v1, fully unrolled over 20 joints, was ~100 KB of SASS and ran at a 51 %
//     instruction-cache hit rate.)  Per-lane joint vectors (line-search origin xold, direction p,
//     previous interpolated solution prev, Jacobian columns) live in shared memory in [element][lane]
//     order (bank = lane: conflict-free) and are walked with running pointers.
//   * joints whose preceding constant rotation is the identity and whose axis is a coordinate axis
//     (every joint of the planar robots, most joints of the arms) take a short path: the joint
//     transform touches two rows / two components only.
//   * the trial point is never stored: x_j = clamp(base_j + t * dir_j), with (base, dir, t) =
//     (xold, p, alam) during a line search and (qa, qb - qa, u) for the first evaluation of an
//     interpolated solve, so starting an instance needs no seed-writing loop.
//   * single precision uses a branch-free Cody-Waite + minimax sincos and MUFU-based reciprocal /
//     rsqrt in the transition logic (IEEE division's slow path was 11 % of all issued instructions).
//
// Restates (reference file:line in the reference repository): NewtonRoot::GlobalSolve as driven by
// IKDB/Klamp't (call sites grr/graph.py:779-790, grr/solver.py:799,821,884).
#pragma once
#include "grr_device.cuh"

namespace grr {

// ---- fast scalar helpers (float: MUFU based; double: exact) ---------------------------------
#ifdef __CUDA_ARCH__
GRR_DEV float fdiv_(float a, float b) { return __fdividef(a, b); }
GRR_DEV float rcp_ge1_(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }  // a >= 1
GRR_DEV float rsqrt_(float a) { return rsqrtf(a); }
GRR_DEV float rsqrt_est_(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }  // estimates only
GRR_DEV int float_bits_(float t) { return __float_as_int(t); }
#else
GRR_DEV float fdiv_(float a, float b) { return a / b; }
GRR_DEV float rcp_ge1_(float a) { return 1.0f / a; }
GRR_DEV float rsqrt_(float a) { return 1.0f / sqrtf(a); }
GRR_DEV float rsqrt_est_(float a) { return 1.0f / sqrtf(a); }
GRR_DEV int float_bits_(float t) { int q; memcpy(&q, &t, 4); return q; }
#endif
GRR_DEV double fdiv_(double a, double b) { return a / b; }
GRR_DEV double rcp_ge1_(double a) { return 1.0 / a; }
#ifdef __CUDA_ARCH__
GRR_DEV double rsqrt_(double a) { return rsqrt(a); }      // (1 / sqrt as an IEEE square root + IEEE division was 5.7 % of the double-precision edge kernel's samples)
#else
GRR_DEV double rsqrt_(double a) { return 1.0 / sqrt(a); }
#endif
GRR_DEV double rsqrt_est_(double a) { return 1.0 / sqrt(a); }

// Running maximum of the step-size tests, max_j |d_j| / max(|x_j|, 1) (NewtonRoot's `test`).  Single precision:
// the value, with a MUFU reciprocal per joint.  Double precision: the maximising fraction (num, den), compared by
// cross-multiplication -- an IEEE double division per joint and test was 3 % of the kernel -- and divided once per pass.
template <typename Real> struct RatioMax;
template <> struct RatioMax<float> {
  float v = 0;
  GRR_DEV void add(float num, float rinv, float) { v = fmaxf(v, num * rinv); }
  GRR_DEV float get() const { return v; }
};
template <> struct RatioMax<double> {
  double n = 0, d = 1;
  GRR_DEV void add(double num, double, double den) { if (num * d > n * den) { n = num; d = den; } }
  GRR_DEV double get() const { return n / d; }
};

// sin/cos for joint angles (|x| << 2^22): Cody-Waite reduction to [-pi/4, pi/4] with the
// round-to-nearest magic constant (no F2I / I2F), cephes minimax polynomials, quadrant fix-up.
GRR_DEV void sincos_joint(float x, float *s, float *c) {
  const float t = fmaf(x, 0.636619772367581343f, 12582912.0f);
  const int q = float_bits_(t);
  const float k = t - 12582912.0f;
  float r = fmaf(k, -1.5707962512969970703f, x);
  r = fmaf(k, -7.5497894158615963534e-08f, r);
  r = fmaf(k, -5.3903029534742383927e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const float ss = (q & 1) ? cp : sp;
  const float cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
// Double precision (float64 build, recheck pass of the exact edge build): the same structure with fdlibm's kernels
// (__kernel_sin / __kernel_cos on [-pi/4, pi/4], < 1 ulp) instead of the library sincos, whose general-purpose code is
// ~170 instructions per call with its 64-bit immediates and slow-path tests -- half of the double-precision chain pass.
GRR_DEV void sincos_kernel_(double r, double *s, double *c) {
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  *s = fma(z * r, fma(z, ps, -1.66666666666666324348e-01), r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  // cos = (1 - qx) - ((z/2 - qx) - z^2 pc) with qx ~ |r|/4 rounded to 24 bits, so that 1 - qx is exact (fdlibm's trick)
  const double qx = (double)(float)(0.25 * fabs(r));
  *c = (1.0 - qx) - ((0.5 * z - qx) - z * (z * pc));
}
GRR_DEV void sincos_joint(double x, double *s, double *c) {
  const double k = rint(x * 6.36619772367581382433e-01);
  // pi/2 in three parts, the first two with 33 significant bits: k * part is exact for |k| < 2^20, i.e. |x| < 1.6e6 rad.
  // (Joint angles live in their limits; only the iterate of a diverging solve of an unbounded joint can leave that
  // range, and such a solve fails in either arithmetic.)
  double r = fma(k, -1.57079632673412561417e+00, x);
  r = fma(k, -6.07710050630396597660e-11, r);
  r = fma(k, -2.02226624879595063154e-21, r);
  double sk, ck;
  sincos_kernel_(r, &sk, &ck);
  const int q = (int)k;
  const double ss = (q & 1) ? ck : sk, cc = (q & 1) ? sk : ck;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
// every joint limit of the chain lies inside [-pi/4, pi/4] (planar 10R / 20R): no reduction, no fix-up
GRR_DEV void sincos_small(float r, float *s, float *c) {
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  *s = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  *c = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
}
GRR_DEV void sincos_small(double x, double *s, double *c) { sincos_kernel_(x, s, c); }

// Reals of shared memory one lane needs: three rotating config buffers (line-search origin xold,
// trial point xcur, previous interpolated solution prev), the direction p and the Jacobian columns.
// MC = rows of a stored Jacobian column: M in general, 2 for the planar-y specialisation (PL)
__host__ __device__ constexpr int col_rows(int M, bool PL) { return PL ? 2 : M; }
template <int MC> __host__ __device__ constexpr int lane_reals(int n) { return n * (4 + MC); }

// Shared-memory view of one lane's joint vectors: element k of an array lives at ptr[k * T].
// The three config buffers change roles by swapping offsets, never by copying.
template <typename Real, int MC>
struct LaneMem {
  Real *base;   // already offset by tid
  int T;        // lanes per CTA (stride)
  int n;
  int o_xold, o_xcur, o_prev;   // element offsets of the three config buffers
  GRR_DEV void init(Real *b, int T_, int n_) {
    base = b; T = T_; n = n_; o_xold = 0; o_xcur = n_ * T_; o_prev = 2 * n_ * T_;
  }
  GRR_DEV Real *xold() const { return base + o_xold; }
  GRR_DEV Real *xcur() const { return base + o_xcur; }
  GRR_DEV Real *prev() const { return base + o_prev; }
  GRR_DEV Real *p() const { return base + 3 * n * T; }
  GRR_DEV Real *cols() const { return base + 4 * n * T; }   // [j][r], r < MC
};

// Where the current trial point comes from: x_j = clamp(a[j*sa] + t * (lerp ? b[j*sb] - a[j*sa] : b[j*sb])).
template <typename Real>
struct XSrc {
  const Real *a, *b;
  int sa, sb;
  Real t;
  bool lerp;
};

// bits of LaneIK::phase.  PH_SEARCH: a line search is running (else: the seed point is being evaluated).
// The other three belong to the near-tie bookkeeping of the FP32 edge phase (FLAG builds of lane_step, see there):
// PH_SOFT: a convergence test fell within the FP32 error of tol and was taken as "not converged"; PH_HARD: a
// decision whose two outcomes cannot both be followed fell within its margin (the pair must be re-tested in
// double precision); PH_INFL: the step length of this solve came out of an interior backtrack (its error is a
// multiple of the plain round-off, margins are widened).
enum : int { PH_SEARCH = 1, PH_SOFT = 2, PH_HARD = 4, PH_INFL = 8,
             // which decision raised PH_HARD (diagnostics; they ride along in the same word at no cost)
             PH_R_CONVX = 1 << 4, PH_R_STPMAX = 1 << 5, PH_R_SLOPE = 1 << 6, PH_R_ALAM = 1 << 7, PH_R_ARMIJO = 1 << 8,
             PH_R_DISC = 1 << 9, PH_R_SOFT2 = 1 << 10, PH_R_SOFTFAIL = 1 << 11, PH_R_MAXIT = 1 << 12, PH_R_LEAF = 1 << 13,
             PH_R_NDIVS = 1 << 14, PH_R_ALL = 0x7ff0,
             PH_PEND = 1 << 15,     // an Armijo near-tie awaits its second derivative sample (see lane_step)
             PH_TIEQ = 1 << 16, PH_TIER = 1 << 17, PH_TSAME = 1 << 18 };   // near-tie request posted / answered / answer's `same` bit

template <typename Real, int M>
struct LaneIK {
  Real fold, slope, alam, alam2, f2, alamin, stpmax, resid;
  Real amb;      // FLAG builds: after IK_OK, radius within which the double-precision solution may lie (0 = round-off only)
  Real t_Dc;             // FLAG builds: answer to a posted near-tie request (directional derivative at the trial point)
  Real t_a, t_D, t_ar;   // FLAG builds: first derivative sample of a pending Armijo near-tie (step length, g.p_c, last clear rejection)
  Real aerr;     // FLAG builds: relative error of the current (interpolated) step length
  Real rad;      // FLAG builds: running estimate of |x(single) - x(double)| over the iterations of this solve (see lane_step)
  int it, evals, phase;
  // the Newton step is folded into the next chain pass.  `first` = that pass derives the
  // direction p_j = -col_j . y from the Jacobian columns still in memory (ny = -y) and stores it.
  bool first, check_convx;
  Real ny[M];
  XSrc<Real> src;
  const Real *res;   // after a final status: the solution, res[j * T]
};

template <typename Real, int M>
GRR_DEV void lane_reset(LaneIK<Real, M> &L) {
  L.it = 0; L.evals = 0; L.phase = 0; L.first = false; L.check_convx = false; 
#pragma unroll
  for (int r = 0; r < M; r++) L.ny[r] = 0;
  L.alam = 0; L.alam2 = 0; L.f2 = 0; L.fold = 0; L.slope = 0; L.alamin = 0; L.stpmax = 0; L.resid = 0; L.amb = 0; L.rad = 0; L.aerr = 0; L.t_a = 0; L.t_D = 0; L.t_ar = 0; L.t_Dc = 0;
}
// Start an instance whose seed is lerp(qa, qb, u) read from (a, b) with strides (sa, sb).
template <typename Real, int M>
GRR_DEV void lane_begin_lerp(LaneIK<Real, M> &L, const Real *a, int sa, const Real *b, int sb, Real u) {
  lane_reset(L);
  L.src.a = a; L.src.b = b; L.src.sa = sa; L.src.sb = sb; L.src.t = u; L.src.lerp = true;
}
// Start an instance whose seed has been written to mem.xold() (t = 0: the direction is ignored).
template <typename Real, int M, int MC>
GRR_DEV void lane_begin_mem(LaneIK<Real, M> &L, const LaneMem<Real, MC> &mem) {
  lane_reset(L);
  L.src.a = mem.xold(); L.src.b = mem.p(); L.src.sa = mem.T; L.src.sb = mem.T; L.src.t = 0; L.src.lerp = false;
  L.res = mem.xold();
}

// short joint: identity preceding rotation, axis = sgn * e_K.  (I1, I2) are the two components the
// rotation mixes: new_I1 = c a_I1 - s a_I2, new_I2 = s a_I1 + c a_I2.
template <typename Real, int M, int K>
GRR_DEV void joint_axis(Real s, Real co, Real sgn, Real *v, Real *G, Real *col) {
  constexpr int I1 = (K + 1) % 3, I2 = (K + 2) % 3;
  // column: G^T (axis x v), axis x v has components I1: -sgn v_I2, I2: +sgn v_I1
  const Real w1 = -sgn * v[I2], w2 = sgn * v[I1];
#pragma unroll
  for (int r = 0; r < 3; r++) col[r] = G[3 * I1 + r] * w1 + G[3 * I2 + r] * w2;
  if (M == 6) {
#pragma unroll
    for (int r = 0; r < 3; r++) col[3 + r] = sgn * G[3 * K + r];
  }
  const Real a1 = v[I1], a2 = v[I2];
  v[I1] = co * a1 - s * a2;
  v[I2] = s * a1 + co * a2;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const Real g1 = G[3 * I1 + r], g2 = G[3 * I2 + r];
    G[3 * I1 + r] = co * g1 - s * g2;
    G[3 * I2 + r] = s * g1 + co * g2;
  }
}

// joints per trip of the double-precision planar pass (2 = the loop single precision uses; see lane_pass_planar_)
#ifndef GRR_F64_PLANAR_UNROLL
#define GRR_F64_PLANAR_UNROLL 5
#endif

// ---- the chain pass: executed by EVERY lane of the warp (uniform joint counter) ----------------
template <typename Real, int M>
struct PassOut {
  Real Fw[M], FL[M], f, fmax, xn2;
  Real s2;                         // |x - xref|^2 when a reference config is given
  Real S[M * (M + 1) / 2];
  Real pn2, test, testx;           // first trial of a Newton step: |p|^2, max |p_j|/max(|x_j|,1), max |dx_j|/max(|x_j|,1)
};

// xref (stride sref): config against which the joint-space distance of the trial point is accumulated
// (the previous interpolated solution of a pair); nullptr = none.
// ---- planar-y specialisation: every joint has identity preceding rotation and axis +-e_y, and the
// goal-link frame equals the last joint frame (the reference's planar_{2,3,5,10,20}R robots).  G stays a
// rotation about y, kept as (cos, sin); columns have two in-plane components (x, z); the normal matrix is
// 2x2.  Same algorithm, a third of the arithmetic.
// KT > 0: the lane count T of the CTA (the stride of the per-lane joint vectors) as a compile-time constant, so that the
// column / trial-point addresses of the two joints of a trip are immediate offsets of one running pointer each
template <typename Real, bool WITH_REF, bool SMALL, bool POS, int KT = 0, typename ChainT>
GRR_DEV void lane_pass_planar_(const ChainT &c, const Target<Real, 3> &tg, const LaneIK<Real, 3> &L,
                                                  const LaneMem<Real, 2> &mem, const Real *xref, int sref,
                                                  PassOut<Real, 3> &o) {
  const int n = c.n;
  const XSrc<Real> &src = L.src;
  // in-plane components: I1 = z (index 2), I2 = x (index 0); Rot_y(t): z' = c z - s x, x' = s z + c x
  Real vz = c.ee[2], vx = c.ee[0], vy = c.ee[1];
  Real gc = 1, gs = 0;                       // G = Rot_y(phi): (cos phi, sin phi)
  Real S11 = 0, S12 = 0, S22 = 0, xn2 = 0, s2 = 0;
  Real pn2 = 0, test = 0, testx = 0;
  RatioMax<Real> rtest, rtestx;           // (chains with joint ranges beyond +-pi/4)
  const int T = KT > 0 ? KT : mem.T;
  const int sa = src.sa, sb = src.sb;
  const Real *pa = src.a + (n - 1) * sa;
  Real *pb = const_cast<Real *>(src.b) + (n - 1) * sb;      // written only by `first` lanes, whose b is mem.p()
  const Real *pr = WITH_REF ? xref + (n - 1) * sref : nullptr;
  Real *pc = mem.cols() + (n - 1) * 2 * T;
  Real *px = mem.xcur() + (n - 1) * T;
  // dir = b - lk * a: the product is exact for lk in {0, 1}, so this equals (lerp ? b - a : b) bit for bit
  const Real lk = src.lerp ? Real(1) : Real(0);
  const Real t = src.t;
  const bool first = L.first;
  const Real ny0 = L.ny[0], ny1 = L.ny[1];
  auto joint = [&](const JointConst<Real> &jc, Real av, Real bv, Real rv, Real *pbw) {
    // first trial of a Newton step: direction from the columns of the accepted point (still in memory
    // at this joint: they are overwritten below), stored as p for the later trials of the line search
    const Real c1o = pc[0], c2o = pc[T], xo = *px;
    const Real dnew = fma_(c2o, ny1, c1o * ny0);
    const Real dir = first ? dnew : fma_(-lk, av, bv);
    if (first) *pbw = dir;
    pn2 = fma_(dir, dir, pn2);
    if (SMALL) {                                        // |x_j| <= pi/4: max(|x_j|, 1) == 1
      test = max_(test, abs_(dir));
      testx = max_(testx, abs_(av - xo));
    } else {
      const Real den = (abs_(av) > Real(1)) ? abs_(av) : Real(1);
      const Real rinv = sizeof(Real) == 4 ? rcp_ge1_(den) : Real(0);
      rtest.add(abs_(dir), rinv, den);
      rtestx.add(abs_(av - xo), rinv, den);
    }
    const Real xj = min_(max_(fma_(t, dir, av), jc.bmin), jc.bmax);
    *px = xj; px -= T;
    xn2 += xj * xj;
    if (WITH_REF) { const Real d = xj - rv; s2 += d * d; }
    Real s, co;
    if (SMALL) sincos_small(xj, &s, &co); else sincos_joint(xj, &s, &co);
    if (!POS) s *= jc.sgn;
    // column = G^T (axis x v): axis x v = sgn (-v_x e_z + v_z e_x); G^T rotates by -phi
    const Real w1 = POS ? -vx : -jc.sgn * vx, w2 = POS ? vz : jc.sgn * vz;
    const Real c1 = gc * w1 + gs * w2, c2 = gc * w2 - gs * w1;
    pc[0] = c1; pc[T] = c2; pc -= 2 * T;
    S11 += c1 * c1; S12 += c1 * c2; S22 += c2 * c2;
    const Real nz = co * vz - s * vx, nx = s * vz + co * vx;
    vz = nz + jc.tp[2]; vx = nx + jc.tp[0]; vy += jc.tp[1];
    const Real ngc = co * gc - s * gs, ngs = s * gc + co * gs;
    gc = ngc; gs = ngs;
  };
  // two joints per trip with two named operand sets: the next joint's constants (an indexed constant-bank
  // load) and shared-memory operands are in flight while the current joint is computed, and no register
  // rotation is needed
  int j = n - 1;
#ifdef __CUDA_ARCH__
  if constexpr (sizeof(Real) == 8 && GRR_F64_PLANAR_UNROLL > 2) {
    // Double precision: the kernels that run this pass (recheck of the exact build, sampling, seeded sweep) hold 4-8
    // warps per SM and are bound by the latency of one joint's dependent chain (the sin / cos polynomials: ~15 dependent
    // DFMAs), not by issue slots.  U joints per trip, operands of all U fetched first: the U polynomial evaluations are
    // independent and overlap; only the short frame update (v, G) is serial.  Same operations per joint in the same
    // order: results are bit-identical to the two-joint loop below.
    constexpr int U = GRR_F64_PLANAR_UNROLL;
#pragma unroll 1
    for (; j >= U - 1; j -= U) {
      JointConst<Real> jcu[U];
      Real au[U], bu[U], ru[U];
      Real *wu[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        jcu[u] = c.jc[j - u];
        au[u] = *pa; bu[u] = *pb; wu[u] = pb; ru[u] = WITH_REF ? *pr : Real(0);
        pa -= sa; pb -= sb;
        if (WITH_REF) pr -= sref;
      }
#pragma unroll
      for (int u = 0; u < U; u++) joint(jcu[u], au[u], bu[u], ru[u], wu[u]);
    }
    // (the joints that are left, fewer than U, run in the two-joint loop below)
  }
#endif
  if (sizeof(Real) == 4 || j >= 0) {
  JointConst<Real> jcA = c.jc[j], jcB;
  Real aA = *pa, bA = *pb, rA = WITH_REF ? *pr : Real(0), aB, bB, rB = 0;
  Real *wA = pb, *wB;
#pragma unroll 1
  for (; j >= 1; j -= 2) {
    jcB = c.jc[j - 1];
    pa -= sa; pb -= sb; aB = *pa; bB = *pb; wB = pb;
    if (WITH_REF) { pr -= sref; rB = *pr; }
    joint(jcA, aA, bA, rA, wA);
    if (j >= 2) {
      jcA = c.jc[j - 2];
      pa -= sa; pb -= sb; aA = *pa; bA = *pb; wA = pb;
      if (WITH_REF) { pr -= sref; rA = *pr; }
    }
    joint(jcB, aB, bB, rB, wB);
  }
  if (j == 0) joint(jcA, aA, bA, rA, wA);
  }
  o.xn2 = xn2; o.s2 = s2;
  if (!SMALL) { test = rtest.get(); testx = rtestx.get(); }
  o.pn2 = pn2; o.test = test; o.testx = testx;
  o.Fw[0] = vx - tg.x[0]; o.Fw[1] = vy - tg.x[1]; o.Fw[2] = vz - tg.x[2];
  o.f = Real(0.5) * (o.Fw[0] * o.Fw[0] + o.Fw[1] * o.Fw[1] + o.Fw[2] * o.Fw[2]);
  // goal-link frame residual, in-plane part: (z, x) rotated by -phi
  o.FL[0] = gc * o.Fw[2] + gs * o.Fw[0];       // component along I1 (z)
  o.FL[1] = gc * o.Fw[0] - gs * o.Fw[2];       // component along I2 (x)
  o.FL[2] = o.Fw[1];
  o.S[0] = S11; o.S[1] = S12; o.S[2] = S22;
  o.fmax = max_(max_(abs_(o.Fw[0]), abs_(o.Fw[1])), abs_(o.Fw[2]));
}
template <typename Real, bool WITH_REF, int KT = 0, typename ChainT>
GRR_DEV void lane_pass_planar(const ChainT &c, const Target<Real, 3> &tg, const LaneIK<Real, 3> &L,
                                                 const LaneMem<Real, 2> &mem, const Real *xref, int sref,
                                                 PassOut<Real, 3> &o) {
  // the shipped planar_{10,20}R robots: limits inside +-pi/4 and every axis +e_y
  if (c.small_angles != 0 && c.axes_positive != 0) lane_pass_planar_<Real, WITH_REF, true, true, KT>(c, tg, L, mem, xref, sref, o);
  else lane_pass_planar_<Real, WITH_REF, false, false, KT>(c, tg, L, mem, xref, sref, o);
}

// The joint loop of the pass.  SPIN: the chain has 'spin' joints (c.spin_mask != 0) -- a compile-time property of the
// kernel (the launchers pick the SPIN instantiations for such chains), so that the kernels every shipped problem runs are
// the code they were before spin joints existed (a run-time test of jc.spin in the common loop, or both loops in one
// kernel, cost the double-precision kernels of the 6-/7-joint arms 4-8 %).
template <typename Real, int M, bool WITH_REF, bool SPIN, int KT = 0, typename ChainT>
GRR_DEV void lane_pass_chain_(const ChainT &c, const LaneIK<Real, M> &L, const LaneMem<Real, M> &mem, const Real *xref, int sref,
                              PassOut<Real, M> &o, Real (&v)[3], Real (&G)[9]) {
  const int n = c.n;
  const XSrc<Real> &src = L.src;
  v[0] = c.ee[0]; v[1] = c.ee[1]; v[2] = c.ee[2];
#pragma unroll
  for (int k = 0; k < 9; k++) G[k] = c.Rtail[k];
#pragma unroll
  for (int k = 0; k < M * (M + 1) / 2; k++) o.S[k] = 0;
  Real xn2 = 0, s2 = 0, pn2 = 0;
  RatioMax<Real> rtest, rtestx;
  const int T = KT > 0 ? KT : mem.T;            // KT > 0: the CTA's lane count as a compile-time constant (immediate offsets)
  const int sa = src.sa, sb = src.sb;
  const Real *pa = src.a + (n - 1) * sa;
  Real *pb = const_cast<Real *>(src.b) + (n - 1) * sb;      // written only by `first` lanes, whose b is mem.p()
  const Real *pr = WITH_REF ? xref + (n - 1) * sref : nullptr;
  Real *pc = mem.cols() + (n - 1) * M * T;
  Real *px = mem.xcur() + (n - 1) * T;
  const Real lk = src.lerp ? Real(1) : Real(0);                // dir = b - lk * a (exact product)
  const Real t = src.t;
  const bool first = L.first;
  // software pipeline: the next joint's constants (an indexed constant-bank load with ~100+ cycles of
  // latency) and operands are fetched while the current joint is computed.  (Two joints per trip with two
  // named operand sets, as in the planar pass, was measured 7-20 % slower here: the body with its four
  // joint kinds then needs more than the 128 / 168 registers the launch bounds allow.)
  JointConst<Real> jn = c.jc[n - 1];
  Real av_n = *pa, bv_n = *pb, rv_n = WITH_REF ? *pr : Real(0);
#pragma unroll 1
  for (int j = n - 1; j >= 0; --j) {
    const JointConst<Real> jc = jn;
    const Real av = av_n, bv = bv_n, rv = rv_n;
    Real *pbw = pb;
    if (j > 0) {
      jn = c.jc[j - 1];
      pa -= sa; pb -= sb;
      av_n = *pa; bv_n = *pb;
      if (WITH_REF) { pr -= sref; rv_n = *pr; }
    }
    // first trial of a Newton step: direction from the columns of the accepted point (still in memory
    // at this joint: they are overwritten below), stored as p for the later trials of the line search
    Real dnew = 0;
#pragma unroll
    for (int r = 0; r < M; r++) dnew = fma_(pc[r * T], L.ny[r], dnew);
    const Real xo = *px;
    const Real dir = first ? dnew : fma_(-lk, av, bv);
    if (first) *pbw = dir;
    pn2 = fma_(dir, dir, pn2);
    const Real den = (abs_(av) > Real(1)) ? abs_(av) : Real(1);
    const Real rinv = sizeof(Real) == 4 ? rcp_ge1_(den) : Real(0);
    rtest.add(abs_(dir), rinv, den);
    rtestx.add(abs_(av - xo), rinv, den);
    Real xj = min_(max_(fma_(t, dir, av), jc.bmin), jc.bmax);
    Real dref = WITH_REF ? xj - rv : Real(0);
    if constexpr (SPIN) {
      if (jc.spin) {
        // 'spin' joint (warp-uniform branch): an interpolated seed lies on the shorter arc a -> b and is normalised to
        // [0, 2 pi) (robot.interpolate); the distance to the previous solution is the wrapped difference (robot.distance);
        // the line search itself runs on the unbounded angle
        if (src.lerp) {
          const Real an = angle_normalize(av);
          xj = angle_normalize(fma_(t, angle_diff(bv, an), an));
        }
        if (WITH_REF) dref = angle_diff(xj, rv);
      }
    }
    *px = xj; px -= T;
    xn2 += xj * xj;
    if (WITH_REF) s2 += dref * dref;
    Real s, co;
    sincos_joint(xj, &s, &co);
    Real col[M];
    if (jc.kind == 0) {
      Real Mj[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Mj[k] = c.A[j][k] + s * c.B[j][k] - co * c.C[j][k];
      const Real a0 = c.ax[j][0], a1 = c.ax[j][1], a2 = c.ax[j][2];
      const Real w0 = a1 * v[2] - a2 * v[1], w1 = a2 * v[0] - a0 * v[2], w2 = a0 * v[1] - a1 * v[0];
#pragma unroll
      for (int r = 0; r < 3; r++) col[r] = G[r] * w0 + G[3 + r] * w1 + G[6 + r] * w2;      // G^T (axis x v)
      if (M == 6) {
#pragma unroll
        for (int r = 0; r < 3; r++) col[3 + r] = G[r] * a0 + G[3 + r] * a1 + G[6 + r] * a2;  // G^T axis
      }
      const Real n0 = Mj[0] * v[0] + Mj[1] * v[1] + Mj[2] * v[2];
      const Real n1 = Mj[3] * v[0] + Mj[4] * v[1] + Mj[5] * v[2];
      const Real n2 = Mj[6] * v[0] + Mj[7] * v[1] + Mj[8] * v[2];
      v[0] = n0; v[1] = n1; v[2] = n2;
      Real nG[9];
      m3_mul(Mj, G, nG);
#pragma unroll
      for (int k = 0; k < 9; k++) G[k] = nG[k];
    } else {
      const Real ss = jc.sgn * s;
      if (jc.kind == 1) joint_axis<Real, M, 0>(ss, co, jc.sgn, v, G, col);
      else if (jc.kind == 2) joint_axis<Real, M, 1>(ss, co, jc.sgn, v, G, col);
      else joint_axis<Real, M, 2>(ss, co, jc.sgn, v, G, col);
    }
    v[0] += jc.tp[0]; v[1] += jc.tp[1]; v[2] += jc.tp[2];
#pragma unroll
    for (int r = 0; r < M; r++) {
      pc[r * T] = col[r];
#pragma unroll
      for (int q = 0; q <= r; q++) o.S[tri(r, q)] += col[r] * col[q];
    }
    pc -= M * T;
  }
  o.xn2 = xn2; o.s2 = s2;
  o.pn2 = pn2; o.test = rtest.get(); o.testx = rtestx.get();
}

template <typename Real, int M, bool WITH_REF, bool SPIN = false, int KT = 0, typename ChainT>
GRR_DEV void lane_pass(const ChainT &c, const Target<Real, M> &tg, const LaneIK<Real, M> &L,
                                          const LaneMem<Real, M> &mem, const Real *xref, int sref, PassOut<Real, M> &o) {
  Real v[3], G[9];
  lane_pass_chain_<Real, M, WITH_REF, SPIN, KT>(c, L, mem, xref, sref, o, v, G);
  o.Fw[0] = v[0] - tg.x[0]; o.Fw[1] = v[1] - tg.x[1]; o.Fw[2] = v[2] - tg.x[2];
  Real f = o.Fw[0] * o.Fw[0] + o.Fw[1] * o.Fw[1] + o.Fw[2] * o.Fw[2];
  m3_Tvec(G, o.Fw, o.FL);
  if (M == 6) {
    Real Re[9], e[3], ew[3];
    m3_Tmul(tg.R, G, Re);
    so3_log(Re, e);
    m3_vec(G, e, ew);
    o.FL[M - 3] = e[0]; o.FL[M - 2] = e[1]; o.FL[M - 1] = e[2];
    o.Fw[M - 3] = ew[0]; o.Fw[M - 2] = ew[1]; o.Fw[M - 1] = ew[2];
    f += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  }
  o.f = Real(0.5) * f;
  Real fmax = 0;
#pragma unroll
  for (int r = 0; r < M; r++) fmax = max_(fmax, abs_(o.Fw[r]));
  o.fmax = fmax;
}

// Cholesky solve of the packed SPD matrix A (lower) with reciprocal square roots.
template <typename Real, int M>
GRR_DEV void chol_solve_fast(Real (&A)[M * (M + 1) / 2], const Real *b, Real *y, Real *inv2sum = nullptr) {
  Real inv[M];
#pragma unroll
  for (int r = 0; r < M; r++) {
#pragma unroll
    for (int s = 0; s <= r; s++) {
      Real acc = A[tri(r, s)];
#pragma unroll
      for (int k = 0; k < s; k++) acc -= A[tri(r, k)] * A[tri(s, k)];
      if (r == s) { inv[r] = rsqrt_(acc); A[tri(r, r)] = acc * inv[r]; }
      else A[tri(r, s)] = acc * inv[s];
    }
  }
#pragma unroll
  for (int r = 0; r < M; r++) {
    Real acc = b[r];
#pragma unroll
    for (int k = 0; k < r; k++) acc -= A[tri(r, k)] * y[k];
    y[r] = acc * inv[r];
  }
#pragma unroll
  for (int r = M - 1; r >= 0; r--) {
    Real acc = y[r];
#pragma unroll
    for (int k = r + 1; k < M; k++) acc -= A[tri(k, r)] * y[k];
    y[r] = acc * inv[r];
  }
  if (inv2sum) {          // sum of the squared diagonal of L^-1: a (lower) estimate of trace(A^-1) ~ 1 / lambda_min
    Real t = 0;
#pragma unroll
    for (int r = 0; r < M; r++) t += inv[r] * inv[r];
    *inv2sum = t;
  }
}

// Numerical Recipes lnsrch backtracking: the next trial step length after a rejected trial
template <bool FLAG, typename Real, int M>
GRR_DEV void lane_backtrack(LaneIK<Real, M> &L, Real f, Real fl_rel, Real fl_f0 = Real(0)) {
  Real tmplam;
  if (L.alam == Real(1)) tmplam = -L.slope / (Real(2) * (f - L.fold - L.slope));
  else {
    const Real rhs1 = f - L.fold - L.alam * L.slope, rhs2 = L.f2 - L.fold - L.alam2 * L.slope;
    const Real a = (rhs1 / (L.alam * L.alam) - rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
    const Real b = (-L.alam2 * rhs1 / (L.alam * L.alam) + L.alam * rhs2 / (L.alam2 * L.alam2)) / (L.alam - L.alam2);
    if (a == Real(0)) tmplam = -L.slope / (Real(2) * b);
    else {
      const Real disc = b * b - Real(3) * a * L.slope;
      if (FLAG && abs_(disc) < fl_rel * b * b) L.phase |= PH_HARD | PH_R_DISC;       // the two branches are not continuous at disc = 0
      if (disc < Real(0)) tmplam = Real(0.5) * L.alam;
      else if (b <= Real(0)) tmplam = (-b + sqrt_(disc)) / (Real(3) * a);
      else tmplam = -L.slope / (b + sqrt_(disc));
    }
    if (tmplam > Real(0.5) * L.alam) tmplam = Real(0.5) * L.alam;
  }
  // min / max clamps are continuous; an interior value carries the cancellation error of f - fold - slope
  if (FLAG) {
    L.aerr = 0;
    if (tmplam > Real(0.1) * L.alam && tmplam < Real(0.5) * L.alam) {
      L.phase |= PH_INFL;
      // f - fold is known to ~2 |F| fl_f; the step length inherits that relative to the differences it is built from
      const Real ef = Real(2) * fl_f0;
      const Real d1 = abs_(f - L.fold - L.alam * L.slope);
      const Real d2 = (L.alam == Real(1)) ? d1 : abs_(L.f2 - L.fold - L.alam2 * L.slope);
      L.aerr = ef / max_(min_(d1, d2), Real(1e-30));
    }
  }
  L.alam2 = L.alam; L.f2 = f;
  L.alam = max_(tmplam, Real(0.1) * L.alam);
  L.src.t = L.alam;
}

// ---- near-ties of the sufficient-decrease test (FLAG builds) -------------------------------------------
// When f - rhs is below the single-precision noise (late in a backtracking sequence, or a step that barely moves)
// the test is decided by the first-order model instead, which single precision resolves.  On the clamped segment
// origin -> trial point, phi(a) = f(clamp(xold + a p)) has phi'(a) = D(a) = g(x(a)) . p_c with p_c = p minus the joints
// that sit on a limit and are pushed outwards, and the exact test at step a reads b(a) = (phi(a) - fold) / a -
// ALF slope <= 0.
//   one sample Dc = D(alam), worst-case curvature bound C |p|^2:  b(a) within D - tgt -+ alam C |p|^2 for every
//     a <= alam, so a clear sign settles this and every shorter step (rejected for good -> the search ends below
//     alamin: stalled; or accepted here);
//   two samples (this trial and the next, shorter one): D is linear in a up to the third-order term, so
//     b(a) = D0 - tgt + a kappa / 2 is settled on (0, amax] by its two ends; amax covers every step length the
//     double-precision search can try after its last clear rejection.
// Anything else flags the pair.
enum : int { TIE_FLAG = 0, TIE_STALL = 1, TIE_ACCEPT = 2, TIE_PENDING = 3 };
// host study only (tools/lane_host.cu defines GRR_TIE_STATS): which branch of the tie logic decided, how often
#if defined(GRR_TIE_STATS) && !defined(__CUDA_ARCH__)
extern long long g_tie_hist[16];
#define GRR_TIE_STAT(i) do { _Pragma("omp atomic") g_tie_hist[i]++; } while (0)
#else
#define GRR_TIE_STAT(i) do { } while (0)
#endif

// z with g_j = col_j . z: the residual in the goal-link frame, rotation rows multiplied by Jl^-T (the Jacobian of the
// log-map residual is Jl^-1 times the stored angular columns, see lane_step)
template <typename Real, int M, bool PL>
GRR_DEV void tie_weights(const PassOut<Real, M> &o, Real *z) {
  constexpr int MC = col_rows(M, PL);
#pragma unroll
  for (int r = 0; r < (MC < 3 ? MC : 3); r++) z[r] = o.FL[r];
  if constexpr (M == 6) {
    Real D[9];
    so3_Jl_inv(&o.FL[3], D);
    m3_Tvec(D, &o.FL[3], &z[3]);
  }
}

// one joint's share of Dc = g . p_c and of the `same` test (see above); base = the lane's shared-memory column
template <typename Real, int MC, typename ChainT>
GRR_DEV void tie_joint(const ChainT &c, const Real *base, int T, int o_xold, int o_xcur, const Real *z, int j, Real &Dc, bool &same) {
  const int n = c.n;
  const Real pj = base[(3 * n + j) * T], a = base[o_xold + j * T], b = base[o_xcur + j * T];
  const Real lo = c.jc[j].bmin, hi = c.jc[j].bmax;
  const bool out_old = (a >= hi && pj > Real(0)) || (a <= lo && pj < Real(0));
  const bool out_new = (b >= hi && pj > Real(0)) || (b <= lo && pj < Real(0));
  if (out_new != out_old) same = false;
  if (!out_old) {
    const Real *pc = base + (4 * n + j * MC) * T;
    Real g = 0;
#pragma unroll
    for (int r = 0; r < MC; r++) g = fma_(pc[r * T], z[r], g);
    Dc = fma_(g, pj, Dc);
  }
}

// The decision once Dc and `same` are known.  Updates the pending first sample (L.t_*, PH_PEND).
template <typename Real, int M, bool PL, typename ChainT>
GRR_DEV int lane_tie_decide(const ChainT &c, LaneIK<Real, M> &L, const PassOut<Real, M> &o, Real Dc, bool same, bool final_sample) {
  const bool pending = (L.phase & PH_PEND) != 0;
  L.phase &= ~PH_PEND;
  if (!same) { GRR_TIE_STAT(0); return TIE_FLAG; }
  Real trS = 0;
  if constexpr (PL) trS = o.S[0] + o.S[2];
  else {
#pragma unroll
    for (int r = 0; r < M; r++) trS += o.S[tri(r, r)];
  }
  const Real ALF = Real(1e-4);
  const Real F0 = sqrt_(Real(2) * L.fold), tgt = ALF * L.slope;
  if (!pending) {
    const Real curv = L.alam * o.pn2 * (trS + F0 * c.fl_curv);
    const Real slack = c.fl_dd * (abs_(Dc) + abs_(L.slope));
    if (Dc - curv - tgt > slack) { GRR_TIE_STAT(6); return TIE_STALL; }
    if (Dc + curv - tgt < -slack) { GRR_TIE_STAT(7); return TIE_ACCEPT; }
    if (!final_sample && Dc - tgt > slack) {               // (D <= tgt: a shorter step cannot be certified as rejected either)
      L.phase |= PH_PEND;
      L.t_a = L.alam; L.t_D = Dc; L.t_ar = (L.alam2 > Real(0)) ? L.alam2 : L.alam;
      GRR_TIE_STAT(8);
      return TIE_PENDING;
    }
    GRR_TIE_STAT(final_sample ? 1 : 2);
    return TIE_FLAG;
  }
  const Real kappa = (L.t_D - Dc) / (L.t_a - L.alam), D0 = Dc - L.alam * kappa;
  const Real amax = max_(L.t_a, Real(0.5) * L.t_ar);
  const Real pn = sqrt_(o.pn2);
  const Real tau = c.fl_curv * o.pn2 * pn * (Real(3) * sqrt_(trS) + F0 * sqrt_(Real(c.n)) * pn);
  const Real s3 = c.fl_dd * (abs_(Dc) + abs_(L.t_D) + abs_(L.slope) + abs_(kappa) * amax) + amax * amax * tau;
  const Real b0 = D0 - tgt, bm = D0 - tgt + Real(0.5) * amax * kappa;
  if (amax * pn < Real(0.05) && b0 > s3 && bm > s3) { GRR_TIE_STAT(9); return TIE_STALL; }
  GRR_TIE_STAT(!(amax * pn < Real(0.05)) ? 3 : (!(b0 > s3) ? 4 : 5));
  return TIE_FLAG;
}

// Near-tie met in lane_step.  Device: the directional derivative is a loop over the joints that would run with one
// lane of the warp active (measured: 5.7 % of all issued instructions); instead the lane posts a request (PH_TIEQ, its
// weights z in L.ny, which is free during a line search), leaves its state untouched and returns; the kernel then
// lets the whole warp compute Dc for it, one joint per lane (warp_tie_service in grr_kernels.cuh), and the lane meets
// the same tie again one trip later -- it re-evaluates the same trial point -- with the answer in L.t_Dc (PH_TIER).
// Host build (one lane): computed on the spot.  Returns false if the request was posted (caller returns IK_RUNNING).
template <typename Real, int M, bool PL, typename ChainT>
GRR_DEV bool lane_tie_operands(const ChainT &c, LaneIK<Real, M> &L, const LaneMem<Real, col_rows(M, PL)> &mem,
                               const PassOut<Real, M> &o, Real &Dc, bool &same) {
  constexpr int MC = col_rows(M, PL);
  if (L.phase & PH_TIER) {
    L.phase &= ~PH_TIER;
    Dc = L.t_Dc; same = (L.phase & PH_TSAME) != 0;
    L.phase &= ~PH_TSAME;
    return true;
  }
#ifdef __CUDA_ARCH__
  Real z[MC];
  tie_weights<Real, M, PL>(o, z);
#pragma unroll
  for (int r = 0; r < MC; r++) L.ny[r] = z[r];
  L.phase |= PH_TIEQ;
  return false;
#else
  Real z[MC];
  tie_weights<Real, M, PL>(o, z);
  Dc = 0; same = true;
  for (int j = 0; j < c.n; j++) tie_joint<Real, MC>(c, mem.base, mem.T, mem.o_xold, mem.o_xcur, z, j, Dc, same);
  return true;
#endif
}

// ---- the transition: line-search bookkeeping; no joint loop.  A new Newton iteration only solves the
// damped normal equations for y and arms `first`; the next chain pass derives p_j = -col_j . y from the
// stored columns, writes it, evaluates the first trial point and returns |p|, the step-size tests and
// the x-convergence test, which are then checked here in the order of the reference algorithm
// (convergence on x, iteration limit, step clipping, descent test) before that pass is used as the
// first line-search evaluation.  slope = g.p is taken in closed form: g = J^T F, p = -J^T y  =>
// g.p = -F^T (J J^T) y = -FL^T S y.  (v4/v5 ran a separate per-joint step loop here: 27-43 instructions
// per joint at 73 % lane efficiency, 19 % of all issued instructions.)
// Returns IK_RUNNING or the final status; after a final status the solution is L.res[j * T].
//
// FLAG (single-precision edge phase): near-tie bookkeeping for bit-identical edge sets.  The reference
// arithmetic is IEEE double; a single-precision run can only end with another verdict if one of its discrete
// decisions is taken within the single-precision error of its threshold (measured: 2e-6 .. 3.4e-5 of the pair
// tests, profiles/r01_parity_report.md).  Every decision is therefore checked against a margin (c.fl_*, filled
// by fill_chain from the handle's flag scale):
//   * convergence tests max|F| < tol within the margin are SOFT: the solve goes on as "not converged" and must
//     converge clearly at the very next test; both outcomes of the tie then lead to "solve succeeded" and the
//     two candidate solutions differ by at most that last step, which widens the margin of the adjacent leaf
//     distance tests (L.amb).  (Ties here are frequent -- every solve crosses tol once -- and almost always
//     harmless, so they are resolved in place: one extra Newton iteration.)
//   * every other decision (Armijo accept, stall alam < alamin, convergence on x, descent test, step clipping,
//     the discriminant branch of the cubic backtrack, a second soft tie or a failure after one) within its
//     margin is HARD: PH_HARD is set, the caller lists the pair and it is re-tested in double precision
//     (k_edges_flat<double, ..., LIST>).
template <typename Real, int M, bool PL, bool FLAG = false, typename ChainT>
GRR_DEV int lane_step(const ChainT &c, LaneIK<Real, M> &L, LaneMem<Real, col_rows(M, PL)> &mem,
                                         const PassOut<Real, M> &o, bool active) {
  const int n = c.n;
  const Real tolf = c.tol, tolx = Real(0.01) * c.tol, ALF = Real(1e-4);
  int st = IK_RUNNING;
  if (!active) return st;
  bool search = true, newstep = false, check_convx = false;
  L.evals++;
  if (L.first) {
    L.first = false;
    if (FLAG && L.check_convx && abs_(o.testx - tolx) < c.fl_x) L.phase |= PH_HARD | PH_R_CONVX;
    if (L.check_convx && o.testx < tolx) st = IK_CONVX;
    else if (L.check_convx && L.it >= c.max_iters) st = IK_MAXIT;
    else {
      Real test = o.test;
      // |p| > stpmax is tested in squares (L.stpmax holds max(|x|^2, n^2), stpmax = 10 sqrt of it): no square root on the
      // common path; the roots are taken only when the step is clipped
      const Real stp2 = Real(100) * L.stpmax;
      if (FLAG && abs_(o.pn2 - stp2) < Real(2) * c.fl_rel * stp2) L.phase |= PH_HARD | PH_R_STPMAX;
      if (o.pn2 > stp2) {
        // clip the step to stpmax (rare): rescale p in memory and evaluate its first trial again
        const Real pn = sqrt_(o.pn2);
        const Real sc = (Real(10) * sqrt_(L.stpmax)) / pn;
        test = 0;
        const int T = mem.T;
        Real *pp = mem.p();
        const Real *px = mem.xold();
#pragma unroll 1
        for (int j = 0; j < n; j++) {
          const Real pj = pp[j * T] * sc;
          pp[j * T] = pj;
          test = max_(test, abs_(pj) / max_(abs_(px[j * T]), Real(1)));
        }
        L.slope *= sc;
        search = false;
      }
      if (FLAG && !(L.slope < -c.fl_rel * L.fold)) L.phase |= PH_HARD | PH_R_SLOPE;
      if (!(L.slope < Real(0))) { L.it++; st = IK_STALL; }
      else L.alamin = fdiv_(tolx, test);
    }
    if (st != IK_RUNNING || !search) L.evals--;       // that pass was not an evaluation of the algorithm
    if (st != IK_RUNNING) L.res = mem.xold();
  }
  if (st == IK_RUNNING && search) {
    L.resid = o.fmax;
    L.res = mem.xcur();
    // convergence test with the soft-tie rule (see above); `conv` is the decision that is followed
    bool conv = o.fmax < tolf;
    auto conv_test = [&]() {
      if (FLAG) {
        // common case first: far from tol and no soft tie pending -> a few instructions.  The margin grows with the
        // iteration count: the iterates of a long (creeping) solve drift apart by more than one evaluation's round-off.
        const Real m1 = c.fl_f * fma_(Real(0.125), Real(L.it), Real(1));
        if (abs_(o.fmax - tolf) < Real(8) * m1 || (L.phase & PH_SOFT)) {
          const bool near = abs_(o.fmax - tolf) < ((L.phase & PH_INFL) ? Real(8) * m1 : m1);
          if (L.phase & PH_SOFT) { if (near || !conv) L.phase |= PH_HARD | PH_R_SOFT2; }
          else if (near) { L.phase |= PH_SOFT; conv = false; }
        }
      }
    };
    if (!(L.phase & PH_SEARCH)) {
      conv_test();
      if (conv) st = IK_OK;
      else if (c.max_iters <= 0) st = IK_MAXIT;
      else { L.stpmax = max_(o.xn2, Real(n) * Real(n)); newstep = true; }       // stpmax = 10 max(|x|, n), kept squared / 100
    } else {
      if (FLAG && abs_(L.alam - L.alamin) < c.fl_alam * L.alamin) L.phase |= PH_HARD | PH_R_ALAM;
      if (L.alam < L.alamin) {
        // the search ends here; a pending near-tie gets its second derivative sample from this trial point
        if (FLAG && (L.phase & PH_PEND)) {
          Real Dc; bool same;
          if (!lane_tie_operands<Real, M, PL>(c, L, mem, o, Dc, same)) { L.evals--; return IK_RUNNING; }
          if (lane_tie_decide<Real, M, PL>(c, L, o, Dc, same, true) != TIE_STALL) L.phase |= PH_HARD | PH_R_ARMIJO;
        }
        L.it++; L.res = mem.xold(); st = IK_STALL;                        // x <- xold
      } else {
        const Real rhs = L.fold + ALF * L.alam * L.slope;
        bool accept = o.f <= rhs;
        if (FLAG) {
          // |f - rhs| against the error of f and fold: |F| dF ~ sqrt(fold) fl_f, compared in squares
          const Real df = o.f - rhs, w = (L.phase & PH_INFL) ? Real(64) * c.fl_f2 : c.fl_f2;
          if (df * df < w * L.fold) {
            Real Dc; bool same;
            if (!lane_tie_operands<Real, M, PL>(c, L, mem, o, Dc, same)) { L.evals--; return IK_RUNNING; }
            const int code = lane_tie_decide<Real, M, PL>(c, L, o, Dc, same, false);
            if (code == TIE_STALL) { L.it++; L.res = mem.xold(); st = IK_STALL; }
            else if (code == TIE_ACCEPT) accept = true;
            else if (code == TIE_PENDING) accept = false;
            else L.phase |= PH_HARD | PH_R_ARMIJO;
          }
        }
        if (st != IK_RUNNING) { }
        else if (!accept) lane_backtrack<FLAG>(L, o.f, c.fl_rel, FLAG ? sqrt_(Real(2) * L.fold) * c.fl_f : Real(0));
        else {
          L.it++;
          if (FLAG) {
            // accepted clear of the margin while an earlier near-tie of this search is still open: that tie stays undecided
            if (L.phase & PH_PEND) { L.phase &= ~PH_PEND; L.phase |= PH_HARD | PH_R_ARMIJO; GRR_TIE_STAT(10); }
            // Drift r of the single-precision iterate from the double-precision one over one accepted step (estimate).
            // A step of amplification A = |step| / |F| (>> 1 near a singular configuration, up to 1 / (2 lambda)) turns
            // the residual round-off fl_f into A fl_f of joint error, carries an existing difference on with the factor
            // 1 + A^2 |F| |dJ| (the step depends on where it is taken), and an interpolated step length adds its own
            // relative error aerr times the step.  (A worst-case bound through |J^+| and |(J J^T + lambda^2)^-1 F| was
            // tried: it holds in all but ~1e-5 of the solves but is astronomically large for 10-30 % of the solves of
            // the 6-/7-joint arms; this estimate is exceeded by the measured drift in ~1e-3 of the solves, by <= 30x.)
            const Real f2x = max_(Real(2) * L.fold, Real(1e-30)), ri = rsqrt_est_(f2x);
            const Real pn = o.pn2 * rsqrt_est_(max_(o.pn2, Real(1e-30)));
            const Real A = L.alam * pn * ri;                              // |step| / |F|
            L.rad = fma_(L.rad, fma_(A * A * (f2x * ri), c.fl_curv1, Real(1)), fma_(L.alam * pn, L.aerr, A * c.fl_f));
          }
          conv_test();
          if (conv) {
            st = IK_OK;
            // a soft tie one test earlier: the double-precision run may have stopped one step before this point
            if (FLAG) {
              L.amb = Real(4) * L.rad;
              if (L.phase & PH_SOFT) L.amb += L.alam * o.pn2 * rsqrt_est_(max_(o.pn2, Real(1e-30)));
            }
          } else { newstep = true; check_convx = true; }
        }
      }
    }
  }
  // (a failure after a soft tie -- the other outcome of the tie was "succeeded" -- and a near-tie still pending at the
  // end of the solve are turned into flags by the caller: pair_note_solve)
  if (newstep) {
    // the accepted point (in xcur) becomes the origin of the next line search: swap roles, no copy
    { const int t = mem.o_xold; mem.o_xold = mem.o_xcur; mem.o_xcur = t; }
    const Real l2 = c.lambda * c.lambda;
    if constexpr (PL) {
      // 2x2 in-plane normal equations (the out-of-plane row has no Jacobian)
      const Real a11 = o.S[0] + l2, a12 = o.S[1], a22 = o.S[2] + l2;
      const Real idet = fdiv_(Real(1), a11 * a22 - a12 * a12);
      const Real y0 = (a22 * o.FL[0] - a12 * o.FL[1]) * idet;
      const Real y1 = (a11 * o.FL[1] - a12 * o.FL[0]) * idet;
      L.ny[0] = -y0; L.ny[1] = -y1;
      L.slope = -(o.FL[0] * (o.S[0] * y0 + o.S[1] * y1) + o.FL[1] * (o.S[1] * y0 + o.S[2] * y1));
    } else {
      Real y[M];
      Real A[M * (M + 1) / 2];
      if (M == 3) {
#pragma unroll
        for (int k = 0; k < 6; k++) A[k] = o.S[k];
#pragma unroll
        for (int r = 0; r < M; r++) A[tri(r, r)] += l2;
        chol_solve_fast<Real, M>(A, o.FL, y);
      } else {
        // rotation rows: the Jacobian of the log-map residual is Jl^-1(e) times the stored angular rows
        Real D[9], Src[9], Srr[9], DSrc[9], DSrr[9];
        so3_Jl_inv(&o.FL[M - 3], D);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int q = 0; q < 3; q++) {
            Src[3 * r + q] = o.S[tri(3 + r, q)];
            Srr[3 * r + q] = (r >= q) ? o.S[tri(3 + r, 3 + q)] : o.S[tri(3 + q, 3 + r)];
          }
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
#pragma unroll
        for (int r = 0; r < M; r++) A[tri(r, r)] += l2;
        chol_solve_fast<Real, M>(A, o.FL, y);
        Real z[3];
        m3_Tvec(D, &y[M - 3], z);
        y[M - 3] = z[0]; y[M - 2] = z[1]; y[M - 1] = z[2];
      }
      // slope = -FL^T S y with the stored-column normal matrix S (packed lower)
      Real slope = 0;
#pragma unroll
      for (int r = 0; r < M; r++) {
        Real sy = 0;
#pragma unroll
        for (int q = 0; q < M; q++) sy += ((r >= q) ? o.S[tri(r, q)] : o.S[tri(q, r)]) * y[q];
        slope += o.FL[r] * sy;
      }
      L.slope = -slope;
#pragma unroll
      for (int r = 0; r < M; r++) L.ny[r] = -y[r];
    }
    L.first = true; L.check_convx = check_convx;
    L.fold = o.f; L.alam = 1; L.alam2 = 0; L.f2 = 0; L.alamin = 0;
    if (FLAG) L.aerr = 0;
    L.src.a = mem.xold(); L.src.b = mem.p(); L.src.sa = mem.T; L.src.sb = mem.T; L.src.t = 1; L.src.lerp = false;
    L.res = mem.xold();
    L.phase |= PH_SEARCH;
  }
  return st;
}

}  // namespace grr
