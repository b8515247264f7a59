// lane_host.cu -- HOST instantiation of the lane machine (csrc/grr_lane.cuh, pair_* of grr_kernels.cuh): the same
// source the kernels run, one pair at a time on CPU threads.  Test / study infrastructure only -- the package never
// loads this library.  Used by tools/flag_study.py to measure, without a GPU, how the single-precision run of
// validEdgeLinear differs from the double-precision run and whether the near-tie flags (FLAG builds of lane_step)
// catch every differing verdict, and by tests/test_lane_host.py to check the host build of the lane machine against
// the oracle.  Host float arithmetic is not bit-identical to the device's (MUFU approximations, FMA contraction), so
// the flag rates measured here are estimates of the device's, confirmed on the GPU by tests/test_gpu_exact.py.
//
// build: nvcc -O2 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fopenmp -shared -o tools/_build/liblane_host.so tools/lane_host.cu
#define GRR_TIE_STATS
#include "../globalredundancyresolution_b200/csrc/grr_kernels.cuh"
#include <omp.h>

namespace grr {
long long g_tie_hist[16] = {0};
void set_error(const char *, ...) {}
template <typename Real, int M, bool PL, bool FLAG>
static void run_pair(const ChainDev<Real, kMaxJ> &c, const Real *qa, const Real *qb, const Real *wa, const Real *wb,
                     Real epsilon, uint8_t *valid, int32_t *info, double *trace = nullptr, int trace_cap = 0) {
  constexpr int MC = col_rows(M, PL);
  Real buf[lane_reals<MC>(kMaxJ)];
  const int n = c.n;
  for (int j = 0; j < lane_reals<MC>(n); j++) buf[j] = 0;
  LaneMem<Real, MC> mem;
  mem.init(buf, 1, n);
  const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));
  const Real thr = eps * Real(c.nlinks_eps);
  const Real thr2 = thr * thr;
  LaneIK<Real, M> L;
  lane_begin_mem(L, mem);
  PairLane<Real> P;
  WInterp<Real, M> W;
  Target<Real, M> tg;
  tg.x[0] = tg.x[1] = tg.x[2] = 0;
  if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
  Real a6[6] = {0, 0, 0, 0, 0, 0}, b6[6] = {0, 0, 0, 0, 0, 0};
  const int dw = (c.mode == 2 ? 6 : 3);
  for (int k = 0; k < dw; k++) { a6[k] = wa[k]; b6[k] = wb[k]; }
  winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
  P.qa = qa; P.qb = qb; P.sa = 1; P.sb = 1;
  int verdict = -1, fail = 0, maxit = 0, fstat = 0, fit = 0; double fres = 0;
  if (!pair_begin<Real, M, FLAG>(c, eps, P, L, c.spin_mask)) verdict = 1;
  else winterp_at<Real, M>(W, P.u_at(1), tg);
  while (verdict < 0) {
    PassOut<Real, M> o;
    const bool first = P.k == 1;
    if constexpr (PL) lane_pass_planar<Real, true>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
    else if (c.spin_mask) lane_pass<Real, M, true, true>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
    else lane_pass<Real, M, true>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
    const int st = lane_step<Real, M, PL, FLAG>(c, L, mem, o, true);
    if (st != IK_RUNNING) {
      P.solves++; P.iters += L.it;
      if (L.it > maxit) maxit = L.it;
      if (trace && P.solves <= trace_cap) {
        // per solve: k, status, iterations, evals, residual, phase bits, leaf s2, target (3), solution (n)
        double *tr = trace + (size_t)(P.solves - 1) * (10 + kMaxJ);
        tr[0] = P.k; tr[1] = st; tr[2] = L.it; tr[3] = (double)L.amb; tr[4] = (double)L.resid; tr[5] = L.phase; tr[6] = (double)o.s2;
        tr[7] = (double)tg.x[0]; tr[8] = (double)tg.x[1]; tr[9] = (double)tg.x[2];
        for (int j = 0; j < n; j++) tr[10 + j] = (double)L.res[j];
      }
      if (FLAG) pair_note_solve(P, L, st);
      if (st != IK_OK) { verdict = 0; fail = 1; fstat = st; fit = L.it; fres = (double)L.resid; }
      else {
        verdict = pair_advance<FLAG>(c, mem, o.s2, thr2, P, L, false, c.spin_mask);
        if (verdict == 0) fail = 2;
        if (verdict < 0) winterp_at<Real, M>(W, P.u_at(P.k), tg);
      }
    }
  }
  *valid = (uint8_t)verdict;
  if (info) { info[0] = P.flag; info[1] = P.ndivs; info[2] = P.solves; info[3] = P.iters; info[4] = fail; info[5] = maxit; info[6] = fstat * 100 + fit; info[7] = (int)fmin(fres * 1e6, 2e9); }
}

template <typename Real, int M, bool FLAG>
static void run_all(const HostChain &h, long long P, const double *qa, const double *qb, const double *wa, const double *wb,
                    double epsilon, uint8_t *valid, int32_t *info) {
  ChainDev<Real, kMaxJ> *c = new ChainDev<Real, kMaxJ>;
  fill_chain(h, *c);
  const bool pl = (M == 3) && is_planar_y(h);
  const int n = h.n, dw = (h.mode == 2 ? 6 : 3);
#pragma omp parallel for schedule(dynamic, 256)
  for (long long t = 0; t < P; t++) {
    Real a[kMaxJ], b[kMaxJ], w0[6], w1[6];
    for (int j = 0; j < n; j++) { a[j] = (Real)qa[t * n + j]; b[j] = (Real)qb[t * n + j]; }
    for (int k = 0; k < dw; k++) { w0[k] = (Real)wa[t * dw + k]; w1[k] = (Real)wb[t * dw + k]; }
    int32_t *inf = info ? info + 8 * t : nullptr;
    if constexpr (M == 3) {
      if (pl) { run_pair<Real, 3, true, FLAG>(*c, a, b, w0, w1, (Real)epsilon, valid + t, inf); continue; }
    }
    run_pair<Real, M, false, FLAG>(*c, a, b, w0, w1, (Real)epsilon, valid + t, inf);
  }
  delete c;
}
}  // namespace grr

using namespace grr;

// 'spin' joint flags applied to the chains of the following lane_host_pairs / lane_host_trace calls (NULL = none)
static int g_spin[GRR_MAX_JOINTS];
extern "C" void lane_host_set_spin(const int32_t *flags, int n) {
  for (int j = 0; j < GRR_MAX_JOINTS; j++) g_spin[j] = (flags && j < n && flags[j]) ? 1 : 0;
}

// precision: 0 = float with near-tie flags, 1 = double, 2 = float without flags.  fl: 10 margins (see grr_fill.cuh) or
// NULL = flag_margins(scale).  info [P][8] = {flag reasons, Ndivs, solves, iters, fail kind, longest solve, status * 100 + iterations and max|F| * 1e6 of the failing solve} (nullable).
extern "C" int lane_host_pairs(const grr_chain_desc *d, const double *fl, double scale, long long P, const double *qa,
                               const double *qb, const double *wa, const double *wb, double epsilon, int precision,
                               uint8_t *valid, int32_t *info) {
  HostChain h;
  memset(&h, 0, sizeof h);
  h.n = d->n; h.mode = d->mode; h.nlinks_eps = d->nlinks_eps; h.max_iters = d->max_iters; h.device = 0;
  h.tol = d->tol; h.lambda = d->lambda;
  memcpy(h.Rp, d->Rp, sizeof(double) * 9 * d->n);
  memcpy(h.tp, d->tp, sizeof(double) * 3 * d->n);
  memcpy(h.axis, d->axis, sizeof(double) * 3 * d->n);
  memcpy(h.qmin, d->qmin, sizeof(double) * d->n);
  memcpy(h.qmax, d->qmax, sizeof(double) * d->n);
  memcpy(h.ee_local, d->ee_local, sizeof h.ee_local);
  memcpy(h.R_tail, d->R_tail, sizeof h.R_tail);
  memcpy(h.R_goal, d->R_goal, sizeof h.R_goal);
  memcpy(h.spin, g_spin, sizeof h.spin);
  flag_margins(h, scale);
  if (fl) memcpy(h.fl, fl, sizeof h.fl);
  const bool m6 = h.mode != GRR_FREE;
  if (precision == 0) { if (m6) run_all<float, 6, true>(h, P, qa, qb, wa, wb, epsilon, valid, info); else run_all<float, 3, true>(h, P, qa, qb, wa, wb, epsilon, valid, info); }
  else if (precision == 1) { if (m6) run_all<double, 6, false>(h, P, qa, qb, wa, wb, epsilon, valid, info); else run_all<double, 3, false>(h, P, qa, qb, wa, wb, epsilon, valid, info); }
  else { if (m6) run_all<float, 6, false>(h, P, qa, qb, wa, wb, epsilon, valid, info); else run_all<float, 3, false>(h, P, qa, qb, wa, wb, epsilon, valid, info); }
  return 0;
}

// one pair with a per-solve trace [cap][10 + 32] (see run_pair); precision as in lane_host_pairs
extern "C" int lane_host_trace(const grr_chain_desc *d, double scale, const double *qa, const double *qb, const double *wa,
                               const double *wb, double epsilon, int precision, double *trace, int cap, int32_t *info) {
  HostChain h;
  memset(&h, 0, sizeof h);
  h.n = d->n; h.mode = d->mode; h.nlinks_eps = d->nlinks_eps; h.max_iters = d->max_iters; h.device = 0;
  h.tol = d->tol; h.lambda = d->lambda;
  memcpy(h.Rp, d->Rp, sizeof(double) * 9 * d->n); memcpy(h.tp, d->tp, sizeof(double) * 3 * d->n);
  memcpy(h.axis, d->axis, sizeof(double) * 3 * d->n); memcpy(h.qmin, d->qmin, sizeof(double) * d->n);
  memcpy(h.qmax, d->qmax, sizeof(double) * d->n); memcpy(h.ee_local, d->ee_local, sizeof h.ee_local);
  memcpy(h.R_tail, d->R_tail, sizeof h.R_tail); memcpy(h.R_goal, d->R_goal, sizeof h.R_goal);
  memcpy(h.spin, g_spin, sizeof h.spin);
  flag_margins(h, scale);
  const bool m6 = h.mode != GRR_FREE, pl = !m6 && is_planar_y(h);
  const int n = h.n, dw = (h.mode == 2 ? 6 : 3);
  uint8_t v = 0;
  auto go = [&](auto real, auto Mtag, auto PLtag, auto FLtag) {
    using Real = decltype(real);
    constexpr int M = decltype(Mtag)::value; constexpr bool PL = decltype(PLtag)::value; constexpr bool FL = decltype(FLtag)::value;
    ChainDev<Real, kMaxJ> *c = new ChainDev<Real, kMaxJ>;
    fill_chain(h, *c);
    Real a[kMaxJ], b[kMaxJ], w0[6], w1[6];
    for (int j = 0; j < n; j++) { a[j] = (Real)qa[j]; b[j] = (Real)qb[j]; }
    for (int k = 0; k < dw; k++) { w0[k] = (Real)wa[k]; w1[k] = (Real)wb[k]; }
    run_pair<Real, M, PL, FL>(*c, a, b, w0, w1, (Real)epsilon, &v, info, trace, cap);
    delete c;
  };
  using T = std::true_type; using F = std::false_type;
  using I3 = std::integral_constant<int, 3>; using I6 = std::integral_constant<int, 6>;
  if (precision == 0) { if (m6) go(0.f, I6{}, F{}, T{}); else if (pl) go(0.f, I3{}, T{}, T{}); else go(0.f, I3{}, F{}, T{}); }
  else { if (m6) go(0.0, I6{}, F{}, F{}); else if (pl) go(0.0, I3{}, T{}, F{}); else go(0.0, I3{}, F{}, F{}); }
  return v;
}

extern "C" void lane_host_margins(const grr_chain_desc *d, double scale, double *fl_out) {
  HostChain h;
  memset(&h, 0, sizeof h);
  h.n = d->n; h.mode = d->mode;
  memcpy(h.tp, d->tp, sizeof(double) * 3 * d->n);
  memcpy(h.ee_local, d->ee_local, sizeof h.ee_local);
  flag_margins(h, scale);
  memcpy(fl_out, h.fl, sizeof h.fl);
}

extern "C" void lane_host_tie_hist(long long *out, int reset) {
  for (int k = 0; k < 16; k++) { out[k] = grr::g_tie_hist[k]; if (reset) grr::g_tie_hist[k] = 0; }
}
