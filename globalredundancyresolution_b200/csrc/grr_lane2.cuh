// grr_lane2.cuh -- two lane machines per thread on the packed FP32 instructions of sm_100 (FFMA2 / FMUL2 / FADD2).
//
// Why: the planar chain pass (grr_lane.cuh, lane_pass_planar_) is bound by instruction ISSUE, not by the FMA pipe
// (profiles/r02_edge_exact_nw1500_ncu_summary.txt: issue slots 77 % busy, FMA pipe 39-42 %): of its ~72 instructions
// per joint 37 are FP32 arithmetic, the rest are shared-memory accesses, address updates, min/max and constant loads.
// A thread that runs TWO independent IK instances (slots) keeps their chain state in float2 registers (.x = slot 0,
// .y = slot 1): every arithmetic instruction becomes one packed instruction for both (same FMA-pipe cycles, half the
// issue slots), the Jacobian columns of the two slots are neighbours in shared memory (one 64-bit access), the joint
// constants and the loop bookkeeping are shared.  What stays per slot: the operand pointers (the two instances are in
// different phases: first evaluation of an interpolated solve reads the staged endpoint configs, a line search reads
// the lane's own xold / p), the role-swapped config buffers, min/max/abs (no packed form) and the whole transition
// logic (lane_step runs once per slot, unchanged).
//
// Specialisation: single precision, planar-y chains whose joint limits lie inside [-pi/4, pi/4] and whose axes are all
// +e_y (ChainDev::small_angles && axes_positive: the reference's planar_10R / planar_20R robots, BASELINE config 2).
// Same algorithm and the same operation order per slot as lane_pass_planar_<float, true, true, true>.
#pragma once
#include "grr_lane.cuh"

namespace grr {

#ifdef __CUDACC__
__device__ __forceinline__ float2 f2_(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2_(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2_(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2_(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2_(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2_(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }   // a - b (b * -1 is exact)
__device__ __forceinline__ float2 neg2_(float2 a) { return make_float2(-a.x, -a.y); }

// sincos_small on both slots
__device__ __forceinline__ void sincos_small2(float2 r, float2 &s, float2 &c) {
  const float2 r2 = mul2_(r, r);
  float2 sp = fma2_(r2, f2_(-1.9515295891e-4f), f2_(8.3321608736e-3f));
  sp = fma2_(sp, r2, f2_(-1.6666654611e-1f));
  s = fma2_(mul2_(sp, r2), r, r);
  float2 cp = fma2_(r2, f2_(2.443315711809948e-5f), f2_(-1.388731625493765e-3f));
  cp = fma2_(cp, r2, f2_(4.166664568298827e-2f));
  c = fma2_(mul2_(cp, r2), r2, fma2_(r2, f2_(-0.5f), f2_(1.0f)));
}

// The chain pass of two slots.  mem[s] are the slots' views of the lane memory: base = lanes + 2 * tid + s,
// T = 2 * blockDim.x, so that element k of slot s lives at lanes[k * T + 2 * tid + s] and the two slots' Jacobian
// columns form one float2.  xref[s] / sref[s]: reference config of the leaf distance (always given here).
template <typename ChainT>
__device__ __forceinline__ void lane_pass_planar2(const ChainT &c, const Target<float, 3> (&tg)[2], const LaneIK<float, 3> (&L)[2],
                                                  const LaneMem<float, 2> (&mem)[2], const float *const (&xref)[2],
                                                  const int (&sref)[2], PassOut<float, 3> (&o)[2]) {
  const int n = c.n;
  int T = mem[0].T;
  asm volatile("" : "+r"(T));       // keep the stride in a register (the compiler otherwise re-derives it from blockDim with a constant load per use)
  float2 vz = f2_(c.ee[2]), vx = f2_(c.ee[0]);
  float vy = c.ee[1];
  float2 gc = f2_(1.0f), gs = f2_(0.0f);
  float2 S11 = f2_(0.0f), S12 = f2_(0.0f), S22 = f2_(0.0f), xn2 = f2_(0.0f), s2 = f2_(0.0f), pn2 = f2_(0.0f);
  float test0 = 0, test1 = 0, testx0 = 0, testx1 = 0;
  const XSrc<float> &s0 = L[0].src, &s1 = L[1].src;
  const int sa0 = s0.sa, sb0 = s0.sb, sa1 = s1.sa, sb1 = s1.sb;
  int sr0 = sref[0], sr1 = sref[1];
  asm volatile("" : "+r"(sr0), "+r"(sr1));
  const float *pa0 = s0.a + (n - 1) * sa0, *pa1 = s1.a + (n - 1) * sa1;
  float *pb0 = const_cast<float *>(s0.b) + (n - 1) * sb0, *pb1 = const_cast<float *>(s1.b) + (n - 1) * sb1;
  const float *pr0 = xref[0] + (n - 1) * sr0, *pr1 = xref[1] + (n - 1) * sr1;
  float *px0 = mem[0].xcur() + (n - 1) * T, *px1 = mem[1].xcur() + (n - 1) * T;
  float2 *pc = reinterpret_cast<float2 *>(mem[0].cols() + (n - 1) * 2 * T);     // slot 0's column element; slot 1's is the next float
  int Th = T >> 1;                                                               // T floats = T/2 float2
  asm volatile("" : "+r"(Th));
  const float2 nlk = f2_(s0.lerp ? -1.0f : 0.0f, s1.lerp ? -1.0f : 0.0f);        // dir = b - lk a
  const float2 t = f2_(s0.t, s1.t);
  const bool first0 = L[0].first, first1 = L[1].first;
  const float2 ny0 = f2_(L[0].ny[0], L[1].ny[0]), ny1 = f2_(L[0].ny[1], L[1].ny[1]);
  auto joint = [&](const JointConst<float> &jc, float2 av, float2 bv, float2 rv, float *pbw0, float *pbw1) {
    const float2 c1o = pc[0], c2o = pc[Th];
    const float2 xo = f2_(*px0, *px1);
    const float2 dnew = fma2_(c2o, ny1, mul2_(c1o, ny0));
    const float2 dlin = fma2_(nlk, av, bv);
    const float2 dir = f2_(first0 ? dnew.x : dlin.x, first1 ? dnew.y : dlin.y);
    if (first0) *pbw0 = dir.x;
    if (first1) *pbw1 = dir.y;
    pn2 = fma2_(dir, dir, pn2);
    const float2 dx = sub2_(av, xo);
    test0 = fmaxf(test0, fabsf(dir.x)); test1 = fmaxf(test1, fabsf(dir.y));          // |x_j| <= pi/4: max(|x_j|, 1) == 1
    testx0 = fmaxf(testx0, fabsf(dx.x)); testx1 = fmaxf(testx1, fabsf(dx.y));
    float2 xj = fma2_(t, dir, av);
    xj.x = fminf(fmaxf(xj.x, jc.bmin), jc.bmax); xj.y = fminf(fmaxf(xj.y, jc.bmin), jc.bmax);
    *px0 = xj.x; *px1 = xj.y; px0 -= T; px1 -= T;
    xn2 = fma2_(xj, xj, xn2);
    const float2 d = sub2_(xj, rv);
    s2 = fma2_(d, d, s2);
    float2 s, co;
    sincos_small2(xj, s, co);
    // column = G^T (axis x v): c1 = gs vz - gc vx, c2 = gc vz + gs vx
    const float2 nvx = neg2_(vx);
    const float2 c1 = fma2_(gc, nvx, mul2_(gs, vz));
    const float2 c2 = fma2_(gs, vx, mul2_(gc, vz));
    pc[0] = c1; pc[Th] = c2; pc -= T;                     // two rows of T floats = T float2 / 2 each
    S11 = fma2_(c1, c1, S11); S12 = fma2_(c1, c2, S12); S22 = fma2_(c2, c2, S22);
    const float2 nz = fma2_(s, nvx, mul2_(co, vz)), nx = fma2_(co, vx, mul2_(s, vz));
    vz = add2_(nz, f2_(jc.tp[2])); vx = add2_(nx, f2_(jc.tp[0])); vy += jc.tp[1];
    const float2 ngc = fma2_(neg2_(s), gs, mul2_(co, gc)), ngs = fma2_(co, gs, mul2_(s, gc));
    gc = ngc; gs = ngs;
  };
  // two joints per trip with two named operand sets (no register rotation): the next joint's constants and
  // shared-memory operands are in flight while the current joint is computed
  int j = n - 1;
  JointConst<float> jcA = c.jc[j], jcB;
  float2 aA = f2_(*pa0, *pa1), bA = f2_(*pb0, *pb1), rA = f2_(*pr0, *pr1), aB, bB, rB;
  float *wA0 = pb0, *wA1 = pb1, *wB0, *wB1;
#pragma unroll 1
  for (; j >= 1; j -= 2) {
    jcB = c.jc[j - 1];
    pa0 -= sa0; pa1 -= sa1; pb0 -= sb0; pb1 -= sb1; pr0 -= sr0; pr1 -= sr1;
    aB = f2_(*pa0, *pa1); bB = f2_(*pb0, *pb1); rB = f2_(*pr0, *pr1); wB0 = pb0; wB1 = pb1;
    joint(jcA, aA, bA, rA, wA0, wA1);
    if (j >= 2) {
      jcA = c.jc[j - 2];
      pa0 -= sa0; pa1 -= sa1; pb0 -= sb0; pb1 -= sb1; pr0 -= sr0; pr1 -= sr1;
      aA = f2_(*pa0, *pa1); bA = f2_(*pb0, *pb1); rA = f2_(*pr0, *pr1); wA0 = pb0; wA1 = pb1;
    }
    joint(jcB, aB, bB, rB, wB0, wB1);
  }
  if (j == 0) joint(jcA, aA, bA, rA, wA0, wA1);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    auto pick = [&](float2 v) { return k == 0 ? v.x : v.y; };
    PassOut<float, 3> &ok = o[k];
    ok.xn2 = pick(xn2); ok.s2 = pick(s2); ok.pn2 = pick(pn2);
    ok.test = k == 0 ? test0 : test1; ok.testx = k == 0 ? testx0 : testx1;
    const float fvx = pick(vx), fvz = pick(vz), fgc = pick(gc), fgs = pick(gs);
    ok.Fw[0] = fvx - tg[k].x[0]; ok.Fw[1] = vy - tg[k].x[1]; ok.Fw[2] = fvz - tg[k].x[2];
    ok.f = 0.5f * (ok.Fw[0] * ok.Fw[0] + ok.Fw[1] * ok.Fw[1] + ok.Fw[2] * ok.Fw[2]);
    ok.FL[0] = fgc * ok.Fw[2] + fgs * ok.Fw[0];
    ok.FL[1] = fgc * ok.Fw[0] - fgs * ok.Fw[2];
    ok.FL[2] = ok.Fw[1];
    ok.S[0] = pick(S11); ok.S[1] = pick(S12); ok.S[2] = pick(S22);
    ok.fmax = fmaxf(fmaxf(fabsf(ok.Fw[0]), fabsf(ok.Fw[1])), fabsf(ok.Fw[2]));
  }
}
#endif  // __CUDACC__

}  // namespace grr
