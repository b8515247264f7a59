// grr_kernels.cuh -- kernels of the roadmap-construction path + their launchers.
// Included by the per-(precision, residual-rows) instantiation units grr_inst_*.cu.
#pragma once
#include "grr_device.cuh"
#include "grr_host.h"
#include "grr_fill.cuh"

namespace grr {

// ------------------------------------------------------------------------------------------
// warp-aggregated counter update
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_add(unsigned long long *ctr, int slot, unsigned v) {
  unsigned tot = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(ctr + slot, (unsigned long long)tot);
}

// ------------------------------------------------------------------------------------------
// S2: random-start IK sampling, one thread per (node, sample) slot
// (parallelQSamplingAtP solver.py:875-885; random phase of sampleConfigurationSpace :817-821)
// ------------------------------------------------------------------------------------------
template <typename Real, int N, int M, int NJ>
__global__ void __launch_bounds__(128)
k_ik_sample(const __grid_constant__ ChainDev<Real, NJ> c, long long Nw, int Nq, int Nqp, const Real *__restrict__ params,
            const int32_t *__restrict__ uid, unsigned long long seed, int sample_offset, Real *__restrict__ q_raw,
            uint8_t *__restrict__ status, uint8_t *__restrict__ iters, unsigned long long *counters) {
  const int n = (N > 0 ? N : c.n);
  const int dw = (c.mode == 2 ? 6 : 3);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < Nw * (long long)Nq;
  unsigned ok = 0, its = 0, evs = 0;
  if (live) {
    const long long node = t / Nq;
    const int s = (int)(t - node * Nq);
    IkLane<Real, N, M> L;
    // random start: Philox(key=seed, ctr=(uid, sample, dof/4, 0)), u = (r>>8) 2^-24, fp32 fma
    const uint32_t id = (uint32_t)uid[node], smp = (uint32_t)(s + sample_offset);
#pragma unroll
    for (int jb = 0; jb < (IkLane<Real, N, M>::NA + 3) / 4; jb++) {
      if (N == 0 && jb * 4 >= n) break;
      uint32_t r[4];
      philox4x32_10(id, smp, (uint32_t)jb, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int j = jb * 4 + k;
        if (j < IkLane<Real, N, M>::NA && (N > 0 || j < n)) {
          float u = (float)(r[k] >> 8) * (1.0f / 16777216.0f);
          L.x[j] = (Real)fmaf(u, c.srange[j], c.slo[j]);
        }
      }
    }
    Target<Real, M> tg;
    Real w[6];
    for (int k = 0; k < dw; k++) w[k] = params[node * dw + k];
    target_from_params<Real, M, NJ>(c, w, tg);
    const int st = ik_solve<Real, N, M, NJ>(c, tg, L);
    Real *dst = q_raw + (node * n) * (long long)Nqp + s;
#pragma unroll
    for (int j = 0; j < IkLane<Real, N, M>::NA; j++) {
      if (N == 0 && j >= n) break;
      dst[(long long)j * Nqp] = L.x[j];
    }
    status[node * Nqp + s] = (uint8_t)st;
    if (iters) iters[node * Nqp + s] = (uint8_t)L.it;
    ok = (st == IK_OK); its = L.it; evs = L.evals;
  }
  if (counters) {
    warp_add(counters, GRR_CNT_IK_SOLVES, live ? 1u : 0u);
    warp_add(counters, GRR_CNT_IK_ITERS, its);
    warp_add(counters, GRR_CNT_IK_EVALS, evs);
    warp_add(counters, GRR_CNT_IK_OK, ok);
  }
}

// ------------------------------------------------------------------------------------------
// V2: seeded IK on explicit (target, seed) instances, SoA q[n][P]
// (RedundancyResolutionGraph.solve, graph.py:779-790)
// ------------------------------------------------------------------------------------------
template <typename Real, int N, int M, int NJ>
__global__ void __launch_bounds__(128)
k_ik_solve_batch(const __grid_constant__ ChainDev<Real, NJ> c, long long P, const Real *__restrict__ targets,
                 Real *__restrict__ q, uint8_t *__restrict__ status, uint8_t *__restrict__ iters,
                 Real *__restrict__ resid, unsigned long long *counters) {
  const int n = (N > 0 ? N : c.n);
  const int dw = (c.mode == 2 ? 6 : 3);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < P;
  unsigned ok = 0, its = 0, evs = 0;
  if (live) {
    IkLane<Real, N, M> L;
#pragma unroll
    for (int j = 0; j < IkLane<Real, N, M>::NA; j++) { if (N == 0 && j >= n) break; L.x[j] = q[(long long)j * P + t]; }
    Target<Real, M> tg;
    Real w[6];
    for (int k = 0; k < dw; k++) w[k] = targets[t * dw + k];
    target_from_params<Real, M, NJ>(c, w, tg);
    const int st = ik_solve<Real, N, M, NJ>(c, tg, L);
#pragma unroll
    for (int j = 0; j < IkLane<Real, N, M>::NA; j++) { if (N == 0 && j >= n) break; q[(long long)j * P + t] = L.x[j]; }
    status[t] = (uint8_t)st;
    if (iters) iters[t] = (uint8_t)L.it;
    if (resid) resid[t] = L.resid;
    ok = (st == IK_OK); its = L.it; evs = L.evals;
  }
  if (counters) {
    warp_add(counters, GRR_CNT_IK_SOLVES, live ? 1u : 0u);
    warp_add(counters, GRR_CNT_IK_ITERS, its);
    warp_add(counters, GRR_CNT_IK_EVALS, evs);
    warp_add(counters, GRR_CNT_IK_OK, ok);
  }
}

// ------------------------------------------------------------------------------------------
// V1: validEdgeLinear for one pair held by one lane (graph.py:938-1006), closed form:
// valid <=> every interpolated IK k=1..Ndivs (seed lerp(qa,qb,u_k), target winterp(wa,wb,u_k))
// converges and every consecutive distance (qa, q_1, ..., q_Ndivs, qb) is <= thr.  The BFS order
// of the reference only changes which failure is met first, never the verdict.
// QA/QB are accessors returning dof j of the two endpoint configs.
// ------------------------------------------------------------------------------------------
struct PairInfo { int ndivs, solves, fail, iters, evals; };

template <typename Real, int N, int M, int NJ, typename QA, typename QB>
__device__ __forceinline__ bool valid_edge_lane(const ChainDev<Real, NJ> &c, const WInterp<Real, M> &W, QA qa, QB qb,
                                                Real eps, Real thr, PairInfo &pi) {
  constexpr int NA = IkLane<Real, N, M>::NA;
  const int n = (N > 0 ? N : c.n);
  Real d2 = 0;
#pragma unroll
  for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; Real d = qa(j) - qb(j); d2 += d * d; }
  const int Ndivs = (int)ceil_(sqrt_(d2) / eps);                 // graph.py:946
  pi.ndivs = Ndivs; pi.solves = 0; pi.fail = 0; pi.iters = 0; pi.evals = 0;
  const Real thr2 = thr * thr;
  Real prev[NA];
#pragma unroll
  for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; prev[j] = qa(j); }
  IkLane<Real, N, M> L;
  for (int k = 1; k <= Ndivs; k++) {
    const Real u = Real(k) / Real(Ndivs + 1);                     // graph.py:961
#pragma unroll
    for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; L.x[j] = qa(j) + u * (qb(j) - qa(j)); }   // :962
    Target<Real, M> tg;
    winterp_at<Real, M>(W, u, tg);
    const int st = ik_solve<Real, N, M, NJ>(c, tg, L);            // :963
    pi.solves++; pi.iters += L.it; pi.evals += L.evals;
    if (st != IK_OK) { pi.fail = 1; return false; }               // :964-966
    Real s2 = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; Real d = L.x[j] - prev[j]; s2 += d * d; prev[j] = L.x[j]; }
    if (s2 > thr2) { pi.fail = 2; return false; }                 // :954-959 (leaf)
  }
  Real s2 = 0;
#pragma unroll
  for (int j = 0; j < NA; j++) { if (N == 0 && j >= n) break; Real d = qb(j) - prev[j]; s2 += d * d; }
  if (s2 > thr2) { pi.fail = 2; return false; }
  return true;
}

// explicit pairs, SoA qa[n][P], qb[n][P]  (validEdge, graph.py:1051-1053)
template <typename Real, int N, int M, int NJ>
__global__ void __launch_bounds__(128)
k_valid_edge_batch(const __grid_constant__ ChainDev<Real, NJ> c, long long P, const Real *__restrict__ qa,
                   const Real *__restrict__ qb, const Real *__restrict__ wa, const Real *__restrict__ wb, Real epsilon,
                   uint8_t *__restrict__ valid, int32_t *__restrict__ info, unsigned long long *counters) {
  const int dw = (c.mode == 2 ? 6 : 3);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < P;
  PairInfo pi = {0, 0, 0, 0, 0};
  bool ok = false;
  if (live) {
    const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));         // graph.py:945
    const Real thr = eps * Real(c.nlinks_eps);                    // graph.py:947
    Real a6[6], b6[6];
    for (int k = 0; k < dw; k++) { a6[k] = wa[t * dw + k]; b6[k] = wb[t * dw + k]; }
    WInterp<Real, M> W;
    winterp_begin<Real, M, NJ>(c, a6, b6, W);
    auto fa = [&](int j) { return qa[(long long)j * P + t]; };
    auto fb = [&](int j) { return qb[(long long)j * P + t]; };
    ok = valid_edge_lane<Real, N, M, NJ>(c, W, fa, fb, eps, thr, pi);
    valid[t] = ok ? 1 : 0;
    if (info) { info[4 * t] = pi.ndivs; info[4 * t + 1] = pi.solves; info[4 * t + 2] = pi.fail; info[4 * t + 3] = pi.iters; }
  }
  if (counters) {
    warp_add(counters, GRR_CNT_PAIRS, live ? 1u : 0u);
    warp_add(counters, GRR_CNT_VALID, ok ? 1u : 0u);
    warp_add(counters, GRR_CNT_IK_SOLVES, (unsigned)pi.solves);
    warp_add(counters, GRR_CNT_IK_ITERS, (unsigned)pi.iters);
    warp_add(counters, GRR_CNT_IK_EVALS, (unsigned)pi.evals);
  }
}

// ------------------------------------------------------------------------------------------
// E1/E3: all-pairs edge construction, one CTA per workspace edge
// (connectConfigurationSpaceOnEdge solver.py:689-697 under connectConfigurationSpace :700-755).
// Both endpoint nodes' config blocks are staged in shared memory; a warp owns one 32-pair strip
// (row a, columns 32w..32w+31) at a time so the verdicts of a strip are one ballot = one mask word.
// ------------------------------------------------------------------------------------------
template <typename Real, int N, int M, int NJ>
__global__ void __launch_bounds__(128)
k_edges_build(const __grid_constant__ ChainDev<Real, NJ> c, const int32_t *__restrict__ wedges,
              const Real *__restrict__ params, const Real *__restrict__ q_kept, const int32_t *__restrict__ count,
              const int32_t *__restrict__ old_count, int Nq, int Nqp, Real epsilon, uint32_t *__restrict__ mask,
              uint32_t *__restrict__ edge_stats, unsigned long long *counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real *sa = reinterpret_cast<Real *>(smem_raw);
  const int n = (N > 0 ? N : c.n);
  Real *sb = sa + n * Nqp;
  __shared__ unsigned s_tested, s_valid;
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
  const int ca = count[iw], cb = count[jw];
  const int oa = old_count ? old_count[iw] : 0, ob = old_count ? old_count[jw] : 0;
  const int dw = (c.mode == 2 ? 6 : 3);
  if (threadIdx.x == 0) { s_tested = 0; s_valid = 0; }
  // stage both node blocks (contiguous n*Nqp chunks)
  {
    const Real *ga = q_kept + (long long)iw * n * Nqp, *gb = q_kept + (long long)jw * n * Nqp;
    for (int k = threadIdx.x; k < n * Nqp; k += blockDim.x) { sa[k] = ga[k]; sb[k] = gb[k]; }
  }
  __syncthreads();
  const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));
  const Real thr = eps * Real(c.nlinks_eps);
  WInterp<Real, M> W;
  {
    Real a6[6], b6[6];
    for (int k = 0; k < dw; k++) { a6[k] = params[(long long)iw * dw + k]; b6[k] = params[(long long)jw * dw + k]; }
    winterp_begin<Real, M, NJ>(c, a6, b6, W);
  }
  const int Wd = (Nq + 31) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  unsigned tested = 0, nvalid = 0, solves = 0, its = 0, evs = 0;
  uint32_t *mrow = mask + e * (long long)Nq * Wd;
  for (int tile = warp; tile < Nq * Wd; tile += nwarps) {
    const int a = tile / Wd, w = tile - a * Wd;
    const int b = w * 32 + lane;
    bool act = (a < ca) && (b < cb) && !(a < oa && b < ob);      // solver.py:694
    bool ok = false;
    if (a < ca && w * 32 < cb) {
      if (act) {
        PairInfo pi;
        auto fa = [&](int j) { return sa[j * Nqp + a]; };
        auto fb = [&](int j) { return sb[j * Nqp + b]; };
        ok = valid_edge_lane<Real, N, M, NJ>(c, W, fa, fb, eps, thr, pi);
        tested++; nvalid += ok; solves += pi.solves; its += pi.iters; evs += pi.evals;
      }
    }
    unsigned word = __ballot_sync(0xffffffffu, ok);
    if (old_count) {   // incremental growth: verdicts of pairs whose endpoints are both old are kept (solver.py:694)
      const unsigned keep = __ballot_sync(0xffffffffu, (a < oa) && (b < ob));
      if (lane == 0) word |= mrow[tile] & keep;
    }
    if (lane == 0) mrow[tile] = word;
  }
  atomicAdd(&s_tested, tested);
  atomicAdd(&s_valid, nvalid);
  if (counters) {
    warp_add(counters, GRR_CNT_PAIRS, tested);
    warp_add(counters, GRR_CNT_VALID, nvalid);
    warp_add(counters, GRR_CNT_IK_SOLVES, solves);
    warp_add(counters, GRR_CNT_IK_ITERS, its);
    warp_add(counters, GRR_CNT_IK_EVALS, evs);
  }
  __syncthreads();
  if (threadIdx.x == 0 && edge_stats) { edge_stats[2 * e] = s_tested; edge_stats[2 * e + 1] = s_valid; }
}

// ------------------------------------------------------------------------------------------
// launchers (one set per <Real, M>; N dispatched at run time over the instantiated list)
// ------------------------------------------------------------------------------------------
inline int check_launch(const char *what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -2; }
  return 0;
}

template <typename Real, int N, int M>
struct Impl {
  static constexpr int NJ = (N > 0 ? N : kMaxJ);
  using Chain = ChainDev<Real, NJ>;

  static int ik_sample(const HostChain &h, long long Nw, int Nq, int Nqp, const void *params, const int32_t *uid,
                       unsigned long long seed, int sample_offset, void *q_raw, uint8_t *status, uint8_t *iters,
                       unsigned long long *counters, cudaStream_t st) {
    Chain c; fill_chain(h, c);
    const long long T = Nw * (long long)Nq;
    if (T == 0) return 0;
    const int bs = 128;
    k_ik_sample<Real, N, M, NJ><<<(unsigned)((T + bs - 1) / bs), bs, 0, st>>>(
        c, Nw, Nq, Nqp, (const Real *)params, uid, seed, sample_offset, (Real *)q_raw, status, iters, counters);
    return check_launch("k_ik_sample");
  }
  static int ik_solve_batch(const HostChain &h, long long P, const void *targets, void *q, uint8_t *status,
                            uint8_t *iters, void *resid, unsigned long long *counters, cudaStream_t st) {
    Chain c; fill_chain(h, c);
    if (P == 0) return 0;
    const int bs = 128;
    k_ik_solve_batch<Real, N, M, NJ><<<(unsigned)((P + bs - 1) / bs), bs, 0, st>>>(
        c, P, (const Real *)targets, (Real *)q, status, iters, (Real *)resid, counters);
    return check_launch("k_ik_solve_batch");
  }
  static int valid_edge_batch(const HostChain &h, long long P, const void *qa, const void *qb, const void *wa,
                              const void *wb, double epsilon, uint8_t *valid, int32_t *info,
                              unsigned long long *counters, cudaStream_t st) {
    Chain c; fill_chain(h, c);
    if (P == 0) return 0;
    const int bs = 128;
    k_valid_edge_batch<Real, N, M, NJ><<<(unsigned)((P + bs - 1) / bs), bs, 0, st>>>(
        c, P, (const Real *)qa, (const Real *)qb, (const Real *)wa, (const Real *)wb, (Real)epsilon, valid, info, counters);
    return check_launch("k_valid_edge_batch");
  }
  static int edges_build(const HostChain &h, long long E, const int32_t *wedges, const void *params, const void *q_kept,
                         const int32_t *count, const int32_t *old_count, int Nq, int Nqp, double epsilon,
                         uint32_t *mask, uint32_t *edge_stats, unsigned long long *counters, cudaStream_t st) {
    Chain c; fill_chain(h, c);
    if (E == 0) return 0;
    const size_t smem = 2 * (size_t)h.n * Nqp * sizeof(Real);
    auto kern = k_edges_build<Real, N, M, NJ>;
    if (smem > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) { set_error("edges_build: shared memory %zu B: %s", smem, cudaGetErrorString(err)); return -2; }
    }
    kern<<<(unsigned)E, 128, smem, st>>>(c, wedges, (const Real *)params, (const Real *)q_kept, count, old_count, Nq, Nqp,
                                         (Real)epsilon, mask, edge_stats, counters);
    return check_launch("k_edges_build");
  }
};

}  // namespace grr
