// grr_kernels.cuh -- kernels of the roadmap-construction path + their launchers (v2: lane machines).
// Included by the per-(precision, residual-rows) instantiation units grr_inst_*.cu.
#pragma once
#include "grr_device.cuh"
#include "grr_lane.cuh"
#include "grr_lane2.cuh"
#include "grr_host.h"
#include "grr_fill.cuh"
#include <stdlib.h>
#include <type_traits>

namespace grr {

// ------------------------------------------------------------------------------------------
// warp-aggregated counter update (all 32 lanes must call)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_add(unsigned long long *ctr, int slot, unsigned v) {
  unsigned tot = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && tot) atomicAdd(ctr + slot, (unsigned long long)tot);
}

struct LaneCounters {
  unsigned solves = 0, iters = 0, evals = 0, ok = 0, pairs = 0, valid = 0;
  __device__ __forceinline__ void flush(unsigned long long *counters) const {
    if (!counters) return;
    warp_add(counters, GRR_CNT_IK_SOLVES, solves);
    warp_add(counters, GRR_CNT_IK_ITERS, iters);
    warp_add(counters, GRR_CNT_IK_EVALS, evals);
    warp_add(counters, GRR_CNT_IK_OK, ok);
    warp_add(counters, GRR_CNT_PAIRS, pairs);
    warp_add(counters, GRR_CNT_VALID, valid);
  }
};

// ------------------------------------------------------------------------------------------
// S2 + V2: IK instances pulled from a global work counter, one lane machine per lane.
//   MODE 0: random-start sampling over (node, sample) slots
//           (parallelQSamplingAtP solver.py:875-885; random phase of sampleConfigurationSpace :817-821)
//   MODE 1: explicit (target, seed) instances, SoA q[n][P]  (RedundancyResolutionGraph.solve, graph.py:779-790)
// ------------------------------------------------------------------------------------------
// double precision, 3 residual rows: CTAs of at most 192 lanes, two per SM (<= 170 registers: 141-160 used, no spills;
// 256 lanes x 2 would cap the build at 128 registers, which the general pass does not fit without spilling).  Steady
// state on one B200: config 3 sampling 12.6 ms (1 CTA of 256 per SM: 21.9 ms), config 2 sampling 8.0 ms.
template <typename Real, int M, int MODE, bool PL, bool SPIN = false>
__global__ void __launch_bounds__((sizeof(Real) == 8 && M == 3) ? 192 : 256, (sizeof(Real) == 8 && M == 3) ? 2 : 1)
k_ik_instances(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long total, int Nq, int Nqp,
               const Real *__restrict__ params, const int32_t *__restrict__ uid, unsigned long long seed, int sample_offset,
               Real *__restrict__ q, uint8_t *__restrict__ status, uint8_t *__restrict__ iters, Real *__restrict__ resid,
               unsigned long long *work, unsigned long long *counters) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(reinterpret_cast<Real *>(smem_raw) + threadIdx.x, (int)blockDim.x, n);
  LaneIK<Real, M> L;
  lane_begin_mem(L, mem);
  Target<Real, M> tg;
  tg.x[0] = tg.x[1] = tg.x[2] = 0;
  if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  LaneCounters cnt;
  long long inst = -1;
  bool done = false;
  for (;;) {
    if (inst < 0 && !done) {
      const long long t = (long long)atomicAdd(work, 1ull);
      if (t >= total) done = true;
      else {
        inst = t;
        Real w[6];
        Real *xo = mem.xold();
        if (MODE == 0) {
          const long long node = t / Nq;
          const int s = (int)(t - node * Nq);
          const uint32_t id = (uint32_t)uid[node], smp = (uint32_t)(s + sample_offset);
          for (int jb = 0; jb * 4 < n; jb++) {
            uint32_t r[4];
            philox4x32_10(id, smp, (uint32_t)jb, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int j = jb * 4 + k;
              if (j < n) {
                const float u = (float)(r[k] >> 8) * (1.0f / 16777216.0f);
                xo[j * mem.T] = (Real)fmaf(u, c.srange[j], c.slo[j]);
              }
            }
          }
          for (int k = 0; k < dw; k++) w[k] = params[node * dw + k];
        } else {
          for (int j = 0; j < n; j++) xo[j * mem.T] = q[(long long)j * total + t];
          for (int k = 0; k < dw; k++) w[k] = params[t * dw + k];
        }
        target_from_params<Real, M, kMaxJ>(c, w, tg);
        lane_begin_mem(L, mem);
      }
    }
    if (__all_sync(0xffffffffu, inst < 0)) break;
    PassOut<Real, M> o;
    if constexpr (PL) lane_pass_planar<Real, false>(c, tg, L, mem, nullptr, 0, o);
    else lane_pass<Real, M, false, SPIN>(c, tg, L, mem, nullptr, 0, o);
    const int st = lane_step<Real, M, PL>(c, L, mem, o, inst >= 0);
    if (inst >= 0) {
      if (st != IK_RUNNING) {
        long long base, stride, sidx;
        if (MODE == 0) {
          const long long node = inst / Nq;
          const int s = (int)(inst - node * Nq);
          base = node * n * (long long)Nqp + s; stride = Nqp; sidx = node * Nqp + s;
        } else { base = inst; stride = total; sidx = inst; }
        for (int j = 0; j < n; j++) q[base + j * stride] = joint_out(kspin, j, L.res[j * mem.T]);
        status[sidx] = (uint8_t)st;
        if (iters) iters[sidx] = (uint8_t)L.it;
        if (resid) resid[sidx] = L.resid;
        cnt.solves++; cnt.iters += L.it; cnt.evals += L.evals; cnt.ok += (st == IK_OK);
        inst = -1;
        lane_begin_mem(L, mem);     // idle lanes keep walking valid memory
      }
    }
  }
  cnt.flush(counters);
}

// ------------------------------------------------------------------------------------------
// S1: the seeded attempts of one wavefront level of sampleConfigurationSpace's sequential sweep
// (solver.py:757-815).  Node i depends on its lower-numbered neighbours only, so all nodes of one level of
// that DAG are processed together.  Instance (l, s), node = node_list[l]: if s < numAdjacent =
// int(|adj|/deg * Nq), solve from config r1 % count[j] of lower neighbour j = adj[r0 % |adj|] (Philox stream 1).
// The random-start attempts of the sweep (solver.py:817-835) do not depend on other nodes at all and are
// computed for the whole graph beforehand by k_ik_instances; k_dedup_sweep then selects the
// Nq - numAdjacent + numFailed of them the sweep would have run.  nadj[l] is written for that kernel.
// ------------------------------------------------------------------------------------------
template <typename Real, int M, bool PL, bool SPIN = false>
__global__ void __launch_bounds__(256)
k_ik_level(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long total, const int32_t *__restrict__ node_list, int Nq,
           int Nqr, const Real *__restrict__ params, const int32_t *__restrict__ uid,
           const int32_t *__restrict__ lower_ptr, const int32_t *__restrict__ lower_idx, const int32_t *__restrict__ degree,
           const Real *__restrict__ Qk, int Nqk, const int32_t *__restrict__ count, unsigned long long seed, int sample_offset,
           int seed_from_adjacent, Real *__restrict__ q_raw, uint8_t *__restrict__ status, int32_t *__restrict__ nadj,
           unsigned long long *work, unsigned long long *counters) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(reinterpret_cast<Real *>(smem_raw) + threadIdx.x, (int)blockDim.x, n);
  LaneIK<Real, M> L;
  lane_begin_mem(L, mem);
  Target<Real, M> tg;
  tg.x[0] = tg.x[1] = tg.x[2] = 0;
  if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  LaneCounters cnt;
  long long li = -1;
  int col = 0;
  bool done = false;
  for (;;) {
    while (li < 0 && !done) {
      const long long t = (long long)atomicAdd(work, 1ull);
      if (t >= total) { done = true; break; }
      const long long l = t / Nq;
      const int s = (int)(t - l * Nq);
      const int node = node_list[l];
      int na = 0;
      for (int k = lower_ptr[node]; k < lower_ptr[node + 1]; k++) na += (count[lower_idx[k]] > 0);   // solver.py:776
      const int deg = degree[node];
      const int numAdjacent = deg > 0 ? (int)((double)na / (double)deg * (double)Nq) : 0;              // :777-778
      if (s == 0) nadj[l] = numAdjacent;
      if (!seed_from_adjacent || s >= numAdjacent) continue;
      uint32_t r[4];
      philox4x32_10((uint32_t)uid[node], (uint32_t)(sample_offset + s), 0u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      int pick = (int)(r[0] % (uint32_t)na), jn = -1;
      for (int k = lower_ptr[node]; k < lower_ptr[node + 1]; k++) {
        const int cand = lower_idx[k];
        if (count[cand] > 0) { if (pick == 0) { jn = cand; break; } pick--; }
      }
      const int cfg = (int)(r[1] % (uint32_t)count[jn]);
      const Real *srcq = Qk + ((long long)jn * n) * Nqk + cfg;
      Real *xo = mem.xold();
      for (int j = 0; j < n; j++) xo[j * mem.T] = srcq[(long long)j * Nqk];
      col = s;
      Real w[6];
      for (int k = 0; k < dw; k++) w[k] = params[(long long)node * dw + k];
      target_from_params<Real, M, kMaxJ>(c, w, tg);
      lane_begin_mem(L, mem);
      li = l;
    }
    if (__all_sync(0xffffffffu, li < 0)) break;
    PassOut<Real, M> o;
    if constexpr (PL) lane_pass_planar<Real, false>(c, tg, L, mem, nullptr, 0, o);
    else lane_pass<Real, M, false, SPIN>(c, tg, L, mem, nullptr, 0, o);
    const int st = lane_step<Real, M, PL>(c, L, mem, o, li >= 0);
    if (li >= 0 && st != IK_RUNNING) {
      Real *dst = q_raw + (li * n) * (long long)Nqr + col;
      for (int j = 0; j < n; j++) dst[(long long)j * Nqr] = joint_out(kspin, j, L.res[j * mem.T]);
      status[li * Nqr + col] = (uint8_t)st;
      cnt.iters += L.it; cnt.evals += L.evals;       // attempts / successes are counted by k_dedup_sweep
      li = -1;
      lane_begin_mem(L, mem);
    }
  }
  cnt.flush(counters);
}

// ------------------------------------------------------------------------------------------
// S1 as ONE launch: the whole sequential sweep of sampleConfigurationSpace (solver.py:773-835) as a dataflow over the
// dependency DAG.  Node i needs the final qlists of its lower-numbered neighbours only; persistent CTAs claim nodes in
// level order (order[] = nodes sorted by level(i) = 1 + max level of the lower neighbours, so everything a claimed node
// waits for was claimed earlier by a CTA that is resident and running), wait on the neighbours' ready flags
// (acquire), run the node's numAdjacent seeded attempts as lane machines, filter seeded + random-start results
// against the node's kept list in the sweep's order, publish the list and release the node's flag.  The random-start
// attempts do not depend on other nodes and are precomputed for the whole graph (k_ik_instances), as in the
// level-by-level version this replaces (280 levels x 2 launches on BASELINE config 2).
// The duplicate filter of a node is done in two steps so that its serial part is short: all candidate-candidate and
// candidate-kept distances at once (bit matrix `close`), then one warp walks the candidates in order.
// Reads of other nodes' count / configs go through L2 (__ldcg): they were written by other SMs during this launch.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_(const int32_t *p) {
  int v;
#ifdef __CUDA_ARCH__
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
  v = *p;
#endif
  return v;
}
__device__ __forceinline__ void st_release_(int32_t *p, int v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *p = v;
#endif
}

// BIASED (solver.py:780-785, `biased=True` on a roadmap that already has configs): a node's number of random-start
// attempts is NConfigsPerPoint * configuration_count / (nodes_with_configs * len(qlist)) with configuration_count the
// RUNNING total of the sweep, so node i also waits for node i - 1 (order[] is then the identity), reads the total node
// i - 1 published and runs its random-start attempts itself, T at a time (they cannot be precomputed: their number is
// known only now), each round filtered into the kept list before the next.
template <typename Real, int M, bool PL, bool BIASED, bool SPIN = false>
__global__ void __launch_bounds__(128)
k_sweep_dataflow(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long Nn, const int32_t *__restrict__ order, int Nq,
                 int Nqr, const Real *__restrict__ params, const int32_t *__restrict__ uid,
                 const int32_t *__restrict__ lower_ptr, const int32_t *__restrict__ lower_idx,
                 const int32_t *__restrict__ degree, Real *Qk, int Nqk, int32_t *count, unsigned long long seed,
                 int sample_offset, int seed_from_adjacent, const Real *__restrict__ q_rand,
                 const uint8_t *__restrict__ st_rand, Real thresh, int32_t *ready, unsigned long long *work,
                 unsigned long long *counters, long long start_total, int nodes_with_configs, long long *total_after) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  // candidates of one filter batch: seeded + random (<= 2 Nq); BIASED: seeded (<= Nq), then rounds of T random ones
  const int NqS = (Nq + 3) & ~3, Cmax = (BIASED && 2 * Nq < 128) ? 128 : 2 * Nq, Wc = (Cmax + 31) >> 5;
  Real *lanes = reinterpret_cast<Real *>(smem_raw);
  Real *cseed = lanes + (size_t)T * lane_reals<col_rows(M, PL)>(n);           // [n][NqS] seeded results
  Real *crand = cseed + (size_t)n * NqS;                                       // BIASED: [n][T] random-start results of one round
  uint32_t *close = reinterpret_cast<uint32_t *>(crand + (BIASED ? (size_t)n * T : 0));   // [Cmax][Wc] candidate s close to candidate t < s
  int32_t *clist = reinterpret_cast<int32_t *>(close + (size_t)Cmax * Wc);     // [Cmax] ok candidates in sweep order: s (seeded) / Nq + s (random)
  int32_t *slot = clist + Cmax;                                                // [Cmax] position in the kept list or -1
  uint8_t *sst = reinterpret_cast<uint8_t *>(slot + Cmax);                     // [max(NqS, T)] status of the attempts of the current batch
  uint8_t *cold = sst + (NqS > T ? NqS : T);                                   // [Cmax] candidate close to an already kept config
  __shared__ long long s_pos, s_total;
  __shared__ int s_nok, s_kept;
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(lanes + tid, T, n);
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  LaneIK<Real, M> L;
  lane_begin_mem(L, mem);
  Target<Real, M> tg;
  tg.x[0] = tg.x[1] = tg.x[2] = 0;
  if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
  LaneCounters cnt;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_pos = (long long)atomicAdd(work, 1ull);
    __syncthreads();
    const long long pos = s_pos;
    if (pos >= Nn) break;
    const int node = order[pos];
    const int lp0 = lower_ptr[node], nlow = lower_ptr[node + 1] - lp0;
    // wait for the lower-numbered neighbours (solver.py:773: the sweep visits nodes in increasing order)
    for (int k = tid; k < nlow; k += T) {
      const int32_t *f = ready + lower_idx[lp0 + k];
      while (ld_acquire_(f) == 0) __nanosleep(64);
    }
    __syncthreads();
    int na = 0;
    for (int k = 0; k < nlow; k++) na += (__ldcg(count + lower_idx[lp0 + k]) > 0);                          // solver.py:776
    const int deg = degree[node];
    const int numAdjRef = deg > 0 ? (int)((double)na / (double)deg * (double)Nq) : 0;                      // :777-778
    const int numAdjacent = seed_from_adjacent ? numAdjRef : 0;                                             // :790 (seeded attempts run)
    Real w[6];
    for (int k = 0; k < dw; k++) w[k] = params[(long long)node * dw + k];
    target_from_params<Real, M, kMaxJ>(c, w, tg);
    const int kept_start = count[node];           // this node's own list: written by this CTA only (or before the launch)
    Real *out = Qk + (long long)node * n * Nqk;
    const Real *randq = q_rand + (long long)node * n * Nqr;
    auto cand_ptr = [&](int ci, int &stride) -> const Real * {
      const int id = clist[ci];
      if (id < Nq) { stride = NqS; return cseed + id; }
      if (BIASED) { stride = T; return crand + (id - Nq); }
      stride = Nqr; return randq + (id - Nq);
    };
    // attempts [0, count) of one batch as lane machines, attempt s on lane s % T; kind 0: seeded from a lower neighbour's
    // config (solver.py:794-799) -> cseed, kind 1 (BIASED): random start number s0 + s (the S2 stream) -> crand
    auto run_attempts = [&](int kind, int nattempt, int s0) {
      int s_mine = tid, s_run = -1;
      for (;;) {
        if (s_run < 0 && s_mine < nattempt) {
          s_run = s_mine; s_mine += T;
          Real *xo = mem.xold();
          if (kind == 0) {
            uint32_t r[4];
            philox4x32_10((uint32_t)uid[node], (uint32_t)(sample_offset + s_run), 0u, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            int pick = (int)(r[0] % (uint32_t)na), jn = -1;
            for (int k = 0; k < nlow; k++) {
              const int cand = lower_idx[lp0 + k];
              if (__ldcg(count + cand) > 0) { if (pick == 0) { jn = cand; break; } pick--; }
            }
            GRR_DCHECK(jn >= 0 && jn < node && na > 0);
            const int cfg = (int)(r[1] % (uint32_t)__ldcg(count + jn));
            GRR_DCHECK(cfg < Nqk);
            const Real *srcq = Qk + ((long long)jn * n) * Nqk + cfg;
            for (int j = 0; j < n; j++) xo[j * mem.T] = __ldcg(srcq + (long long)j * Nqk);
          } else {
            const uint32_t id = (uint32_t)uid[node], smp = (uint32_t)(s0 + s_run + sample_offset);
            for (int jb = 0; jb * 4 < n; jb++) {
              uint32_t r[4];
              philox4x32_10(id, smp, (uint32_t)jb, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int j = jb * 4 + k;
                if (j < n) xo[j * mem.T] = (Real)fmaf((float)(r[k] >> 8) * (1.0f / 16777216.0f), c.srange[j], c.slo[j]);
              }
            }
          }
          lane_begin_mem(L, mem);
        }
        if (__all_sync(0xffffffffu, s_run < 0)) break;
        PassOut<Real, M> o;
        if constexpr (PL) lane_pass_planar<Real, false>(c, tg, L, mem, nullptr, 0, o);
        else lane_pass<Real, M, false, SPIN>(c, tg, L, mem, nullptr, 0, o);
        const int st = lane_step<Real, M, PL>(c, L, mem, o, s_run >= 0);
        if (s_run >= 0 && st != IK_RUNNING) {
          Real *dst = (kind == 0) ? cseed + s_run : crand + s_run;
          const int ds = (kind == 0) ? NqS : T;
          for (int j = 0; j < n; j++) dst[j * ds] = joint_out(kspin, j, L.res[j * mem.T]);
          sst[s_run] = (uint8_t)st;
          cnt.iters += L.it; cnt.evals += L.evals;
          s_run = -1;
          lane_begin_mem(L, mem);
        }
      }
    };
    // filter the candidates clist[0, C) (in order) against the kept list out[0, kept_in) and against each other, append
    // the survivors (first wins, solver.py:802-813,824-835); s_kept = new length.  All distances at once into a bit
    // matrix, then one warp walks the candidates in order.
    auto filter_batch = [&](int C, int kept_in) {
      for (int k = tid; k < C * Wc; k += T) close[k] = 0;
      for (int k = tid; k < C; k += T) cold[k] = 0;
      __syncthreads();
      const int npair = C * (C - 1) / 2;
      for (int pidx = tid; pidx < npair + C * kept_in; pidx += T) {
        int sidx, tidx;
        const bool vs_old = pidx >= npair;
        if (!vs_old) {
          // row s has s entries: s = floor((1 + sqrt(1 + 8 p)) / 2)
          sidx = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pidx)) * 0.5f);
          while (sidx * (sidx - 1) / 2 > pidx) sidx--;
          while ((sidx + 1) * sidx / 2 <= pidx) sidx++;
          tidx = pidx - sidx * (sidx - 1) / 2;
        } else { sidx = (pidx - npair) / kept_in; tidx = (pidx - npair) - sidx * kept_in; }
        int ss, ts;
        const Real *a = cand_ptr(sidx, ss);
        Real d2 = 0;
        if (!vs_old) {
          const Real *b = cand_ptr(tidx, ts);
          for (int j = 0; j < n; j++) { Real d = joint_diff(kspin, j, a[(long long)j * ss], b[(long long)j * ts]); d2 += d * d; }
          if (sqrt_(d2) < thresh) atomicOr(&close[sidx * Wc + (tidx >> 5)], 1u << (tidx & 31));
        } else {
          for (int j = 0; j < n; j++) { Real d = joint_diff(kspin, j, a[(long long)j * ss], out[j * Nqk + tidx]); d2 += d * d; }
          if (sqrt_(d2) < thresh) cold[sidx] = 1;
        }
      }
      __syncthreads();
      if (tid < 32) {
        uint32_t keepw[4] = {0, 0, 0, 0};        // lane owns words lane, lane + 32, ... of the keep mask (Wc <= 128)
        int kept = kept_in;
        for (int sidx = 0; sidx < C; sidx++) {
          bool hit = false;
#pragma unroll
          for (int q = 0; q < 4; q++) { const int wd = lane + 32 * q; if (wd < Wc) hit |= (close[sidx * Wc + wd] & keepw[q]) != 0; }
          hit = __any_sync(0xffffffffu, hit) || cold[sidx];
          const bool take = !hit && kept < Nqk;
          if (take) {
            const int wd = sidx >> 5;
#pragma unroll
            for (int q = 0; q < 4; q++) if (wd == lane + 32 * q) keepw[q] |= 1u << (sidx & 31);
          }
          if (lane == 0) slot[sidx] = take ? kept : -1;
          kept += take;
        }
        if (lane == 0) s_kept = kept;
      }
      __syncthreads();
      for (int k = tid; k < C * n; k += T) {
        const int sidx = k / n, j = k - sidx * n;
        const int sl = slot[sidx];
        GRR_DCHECK(sl < Nqk);
        if (sl >= 0) { int ss; const Real *a = cand_ptr(sidx, ss); out[j * Nqk + sl] = a[(long long)j * ss]; }
      }
      __syncthreads();
    };
    // ---- seeded attempts (solver.py:786-815) ----
    run_attempts(0, numAdjacent, 0);
    __syncthreads();
    int nfail = 0;
    for (int s = 0; s < numAdjacent; s++) nfail += (sst[s] != 0);          // (every thread; numAdjacent <= Nq)
    if constexpr (!BIASED) {
      // candidates in the sweep's order: seeded, then the first Nq - numAdjacent + numFailed precomputed random-start results
      const int nseed = numAdjacent;
      const int nrand = Nq - numAdjRef + nfail;                            // solver.py:779,817
      if (tid < 32) {
        const uint8_t *str = st_rand + (long long)node * Nqr;
        int nok = 0;
        for (int base = 0; base < nseed + nrand; base += 32) {
          const int ci = base + lane;
          bool ok = false;
          if (ci < nseed) ok = (sst[ci] == 0);
          else if (ci < nseed + nrand) ok = (str[ci - nseed] == 0);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          GRR_DCHECK(nok + __popc(m) <= Cmax && nseed + nrand <= Cmax && nrand <= Nqr);
          if (ok) clist[nok + __popc(m & ((1u << lane) - 1u))] = (ci < nseed) ? ci : Nq + (ci - nseed);
          nok += __popc(m);
        }
        if (lane == 0) {
          s_nok = nok;
          if (counters) {
            atomicAdd(counters + GRR_CNT_IK_SOLVES, (unsigned long long)(nseed + nrand));
            atomicAdd(counters + GRR_CNT_IK_OK, (unsigned long long)nok);
          }
        }
      }
      __syncthreads();
      filter_batch(s_nok, kept_start);
    } else {
      // seeded results first
      if (tid < 32) {
        int nok = 0;
        for (int base = 0; base < numAdjacent; base += 32) {
          const int ci = base + lane;
          const bool ok = ci < numAdjacent && sst[ci] == 0;
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) clist[nok + __popc(m & ((1u << lane) - 1u))] = ci;
          nok += __popc(m);
        }
        if (lane == 0) s_nok = nok;
      }
      __syncthreads();
      unsigned long long n_ok = (unsigned long long)s_nok;
      filter_batch(s_nok, kept_start);
      // the running total of the sweep: published by node - 1 (solver.py:763,784,812,833)
      if (tid == 0) {
        long long tot = start_total;
        if (node > 0) { while (ld_acquire_(ready + node - 1) == 0) __nanosleep(64); tot = *((volatile long long *)total_after + (node - 1)); }
        s_total = tot;
      }
      __syncthreads();
      // numRandomSamples = 0 for a node without configs, else NConfigsPerPoint * configuration_count / (nodes_with_configs * len(qlist))
      // (Python 2 integer division) with the total and the node's length as they were when its turn began
      const long long conf = s_total;            // (:779-785 come before this node's own attempts)
      long long nrand_ll = (kept_start == 0 || nodes_with_configs == 0) ? 0
                         : ((long long)Nq * conf) / ((long long)nodes_with_configs * (long long)kept_start);
      const long long nrand = nrand_ll + nfail;
      for (long long s0 = 0; s0 < nrand; s0 += T) {
        const int nb = (int)((nrand - s0 < T) ? (nrand - s0) : T);
        run_attempts(1, nb, (int)s0);
        __syncthreads();
        if (tid < 32) {
          int nok = 0;
          for (int base = 0; base < nb; base += 32) {
            const int ci = base + lane;
            const bool ok = ci < nb && sst[ci] == 0;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) clist[nok + __popc(m & ((1u << lane) - 1u))] = Nq + ci;
            nok += __popc(m);
          }
          if (lane == 0) s_nok = nok;
        }
        __syncthreads();
        n_ok += (unsigned long long)s_nok;
        filter_batch(s_nok, s_kept);
      }
      if (tid == 0) {
        total_after[node] = s_total + (long long)(s_kept - kept_start);
        if (counters) {
          atomicAdd(counters + GRR_CNT_IK_SOLVES, (unsigned long long)numAdjacent + (unsigned long long)nrand);
          atomicAdd(counters + GRR_CNT_IK_OK, n_ok);
        }
      }
    }
    if (tid == 0) {
      count[node] = s_kept;
      __threadfence();
      st_release_(ready + node, 1);
    }
  }
  cnt.flush(counters);
}

// ------------------------------------------------------------------------------------------
// V1 state of one pair held by one lane: validEdgeLinear (graph.py:938-1006) in closed form:
// valid <=> every interpolated IK k=1..Ndivs (seed lerp(qa,qb,u_k), target winterp(wa,wb,u_k))
// converges and every consecutive distance (qa, q_1, ..., q_Ndivs, qb) is <= thr.  The BFS order of
// the reference only changes which failure is met first, never the verdict.
// The endpoint configs are read in place: qa[j*sa], qb[j*sb].
// ------------------------------------------------------------------------------------------
template <typename Real>
struct PairLane {
  const Real *qa, *qb;
  int sa, sb;
  int k, ndivs;       // current interpolation index, graph.py:946
  int solves, iters;  // diagnostics
  // FLAG builds (single-precision edge phase with near-tie bookkeeping, see lane_step): flag != 0 <=> some decision
  // of this pair fell within the single-precision error of its threshold and the pair is re-tested in double
  // precision; amb_prev = ambiguity radius (LaneIK::amb) of the previous interpolated solution.
  int flag;
  int k0;             // FLAG builds: first interpolation index with a near-tie (0: Ndivs itself); recheck kernel: restart index
  Real amb_prev;
  Real inv;           // single precision: 1 / (Ndivs + 1), so that u_k = k * inv costs no division per solve
  // u_k = k / (Ndivs + 1) (graph.py:961): the exact quotient in double precision (the reference's value), k * inv in
  // single precision (within one ulp of it; the IEEE division at every solve boundary was 3 % of the stall samples)
  GRR_DEV Real u_at(int kk) const {
    if constexpr (sizeof(Real) == 4) return Real(kk) * inv;
    else return Real(kk) / Real(ndivs + 1);
  }
};

// Ndivs = ceil(robot.distance(qa, qb) / eps) (graph.py:944-946).  FLAG: bit 30 of the result is set when dist / eps
// lies within fl_nd of an integer (the rounding may go the other way in double precision).
constexpr int kNdFlag = 1 << 30;
template <typename Real, bool FLAG = false>
GRR_DEV int pair_ndivs(int n, unsigned spin, Real eps, const Real *qa, int sa, const Real *qb, int sb, Real fl_nd = Real(0)) {
  Real d2 = 0;
  if (spin) for (int j = 0; j < n; j++) { const Real d = joint_diff(spin, j, qa[j * sa], qb[j * sb]); d2 += d * d; }
  else for (int j = 0; j < n; j++) { const Real d = qa[j * sa] - qb[j * sb]; d2 += d * d; }
  const Real r = sqrt_(d2) / eps, cr = ceil_(r);
  int nd = (int)cr;
  if (FLAG && (cr - r < fl_nd || r - (cr - Real(1)) < fl_nd)) nd |= kNdFlag;
  return nd;
}

// Arm the seed of k=1 for a pair whose Ndivs is known.  Returns false if the pair is decided already.
template <typename Real, int M>
GRR_DEV bool pair_arm(int ndivs, PairLane<Real> &P, LaneIK<Real, M> &L) {
  P.flag = (ndivs & kNdFlag) ? PH_R_NDIVS : 0; P.k0 = 0; P.amb_prev = 0;
  P.ndivs = ndivs & ~kNdFlag;
  P.k = 1; P.solves = 0; P.iters = 0;
  if (P.ndivs <= 0) return false;                                // dist == 0: single leaf of length 0 -> valid
  P.inv = Real(1) / Real(P.ndivs + 1);
  lane_begin_lerp<Real, M>(L, P.qa, P.sa, P.qb, P.sb, P.u_at(1));                     // graph.py:961-962
  return true;
}

// Begin pair: distance -> Ndivs, arm the seed of k=1.  Returns false if the pair is decided already.
template <typename Real, int M, bool FLAG = false, typename ChainT>
GRR_DEV bool pair_begin(const ChainT &c, Real eps, PairLane<Real> &P, LaneIK<Real, M> &L, unsigned spin = 0u) {
  return pair_arm<Real, M>(pair_ndivs<Real, FLAG>(c.n, spin, eps, P.qa, P.sa, P.qb, P.sb, c.fl_nd), P, L);
}

// One interpolated solve finished with IK_OK; s2 = |q_k - q_{k-1}|^2 was accumulated by the final
// pass.  Leaf test, then either the closing leaf (q_Ndivs, qb) or the next k: the solution buffer
// becomes `prev` by swapping roles.  Returns +1 valid, 0 invalid, -1 continue (seed armed).
// FLAG: a leaf distance within its margin of the threshold (round-off fl_leaf plus the ambiguity radii of the two
// solutions it joins) flags the pair.
// FLAG bookkeeping at the end of every interpolated solve (converged or not): collect the solve's near-tie reasons.
template <typename Real, int M>
GRR_DEV void pair_note_solve(PairLane<Real> &P, const LaneIK<Real, M> &L, int st) {
  int r = L.phase & PH_R_ALL;
  if (st > IK_OK && (L.phase & PH_SOFT)) r |= PH_R_SOFTFAIL;     // the other outcome of the soft tie was "solve succeeded"
  if (L.phase & PH_PEND) r |= PH_R_ARMIJO;                       // a near-tie that never got its second sample
  if (r && !P.flag) P.k0 = P.k;
  P.flag |= r;
}

// skip_first_leaf: the recheck kernel restarts a pair at solve kstart > 1, whose predecessor has not been computed:
// the leaf (kstart - 1, kstart) was settled by the single-precision pass.
// Closing leaf (q_Ndivs, qb) of a pair once its squared length e2 is known.
template <bool FLAG, typename Real, int M, typename ChainT>
GRR_DEV int pair_close(const ChainT &c, Real e2, Real thr2, PairLane<Real> &P, const LaneIK<Real, M> &L) {
  if (FLAG) {
    const Real thr = thr2 * rsqrt_est_(thr2);
    const Real amb = min_(L.amb, Real(0.5) * thr);
    if (abs_(e2 - thr2) < thr2 * (Real(2) * c.fl_leaf) + amb * (Real(2) * thr + amb)) { if (!P.flag) P.k0 = P.k; P.flag |= PH_R_LEAF; }
  }
  return (e2 > thr2) ? 0 : 1;
}
constexpr int kNeedClose = -2;     // pair_advance<.., COOP>: the closing leaf is left to the warp (warp_close_service)

// COOP: instead of summing the closing leaf's squared length joint by joint on this lane alone (a loop that ran with
// 1.5 of 32 lanes active: 2.6 % of the dense kernel's issue slots), return kNeedClose; the kernel lets the whole warp
// compute it, one joint per lane.
template <bool FLAG = false, bool COOP = false, typename Real, int M, int MC, typename ChainT>
GRR_DEV int pair_advance(const ChainT &c, LaneMem<Real, MC> &mem, Real s2, Real thr2, PairLane<Real> &P,
                         LaneIK<Real, M> &L, bool skip_first_leaf = false, unsigned spin = 0u) {
  Real thr = 0;
  if (FLAG) {
    // margin of the leaf tests: round-off plus the ambiguity radii of the two solutions, the latter capped at half the
    // threshold (the drift estimate of a sensitive solve is astronomically large: such a solve then flags its leaves
    // only when their length lies within 0.5 .. 1.5 thr)
    thr = thr2 * rsqrt_est_(thr2);
    const Real amb = min_(L.amb + P.amb_prev, Real(0.5) * thr);
    if (abs_(s2 - thr2) < thr2 * (Real(2) * c.fl_leaf) + amb * (Real(2) * thr + amb)) { if (!P.flag) P.k0 = P.k; P.flag |= PH_R_LEAF; }
    P.amb_prev = L.amb;
  }
  if (!skip_first_leaf && s2 > thr2) return 0;                   // graph.py:954-959 (leaf)
  if (P.k == P.ndivs) {
    if (COOP) return kNeedClose;
    const int n = c.n, T = mem.T;
    Real e2 = 0;
    if (spin) for (int j = 0; j < n; j++) { const Real e = joint_diff(spin, j, P.qb[j * P.sb], L.res[j * T]); e2 += e * e; }
    else for (int j = 0; j < n; j++) { const Real e = P.qb[j * P.sb] - L.res[j * T]; e2 += e * e; }
    return pair_close<FLAG>(c, e2, thr2, P, L);
  }
  { const int t = mem.o_prev; mem.o_prev = mem.o_xcur; mem.o_xcur = t; }   // IK_OK leaves the solution in xcur
  P.k++;
  lane_begin_lerp<Real, M>(L, P.qa, P.sa, P.qb, P.sb, P.u_at(P.k));
  return -1;
}

// Serves the near-tie requests of a warp's lanes (lane_tie_operands): for each requesting lane the whole warp computes
// the directional derivative Dc = sum over the free joints of (col_j . z) p_j and the `same` test, one joint per lane,
// reading the requester's shared-memory column.  All 32 lanes must call.
// LS = shared-memory columns per thread (1; 2 in the two-slot kernel, where a slot's columns of neighbouring threads are 2 apart)
template <typename Real, int M, bool PL, int LS = 1, typename ChainT>
__device__ __forceinline__ void warp_tie_service(const ChainT &c, LaneIK<Real, M> &L, const LaneMem<Real, col_rows(M, PL)> &mem,
                                                 bool running) {
#ifdef __CUDA_ARCH__
  constexpr int MC = col_rows(M, PL);
  unsigned tq = __ballot_sync(0xffffffffu, running && (L.phase & PH_TIEQ));
  const int lane = threadIdx.x & 31;
  while (tq) {
    const int src = __ffs(tq) - 1;
    tq &= tq - 1;
    Real z[MC];
#pragma unroll
    for (int r = 0; r < MC; r++) z[r] = __shfl_sync(0xffffffffu, L.ny[r], src);
    const int oxo = __shfl_sync(0xffffffffu, mem.o_xold, src), oxc = __shfl_sync(0xffffffffu, mem.o_xcur, src);
    const Real *base = mem.base + LS * (src - lane);       // lanes of a warp are adjacent columns of the [element][lane] layout
    Real part = 0;
    bool same = true;
    for (int j = lane; j < c.n; j += 32) tie_joint<Real, MC>(c, base, mem.T, oxo, oxc, z, j, part, same);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const bool all_same = __all_sync(0xffffffffu, same);
    if (lane == src) {
      L.t_Dc = part;
      L.phase = (L.phase & ~(PH_TIEQ | PH_TSAME)) | PH_TIER | (all_same ? PH_TSAME : 0);
    }
  }
#endif
}

// Serves the closing-leaf requests of a warp's lanes (pair_advance<.., COOP>): for each requesting lane the whole warp sums
// |qb - q_Ndivs|^2, one joint per lane (n <= 32), reading the requester's staged endpoint config and its solution in
// shared memory.  Returns this lane's e2 (meaningful on requesting lanes).  All 32 lanes must call.
template <typename Real, int M>
__device__ __forceinline__ Real warp_close_service(int n, unsigned spin, int T, const PairLane<Real> &P, const LaneIK<Real, M> &L, bool request) {
  Real mine = 0;
#ifdef __CUDA_ARCH__
  unsigned rq = __ballot_sync(0xffffffffu, request);
  const int lane = threadIdx.x & 31;
  while (rq) {
    const int src = __ffs(rq) - 1;
    rq &= rq - 1;
    const Real *qb = reinterpret_cast<const Real *>(__shfl_sync(0xffffffffu, (unsigned long long)P.qb, src));
    const Real *rs = reinterpret_cast<const Real *>(__shfl_sync(0xffffffffu, (unsigned long long)L.res, src));
    const int sb = __shfl_sync(0xffffffffu, P.sb, src);
    Real e = 0;
    if (lane < n) { e = joint_diff(spin, lane, qb[lane * sb], rs[lane * T]); e *= e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == src) mine = e;
  }
#endif
  return mine;
}

// Near-tie list of the exact edge build: word 0 = entries appended (may exceed the capacity: the excess is dropped and
// reported), entries from word 2 on, two words each: {workspace edge << 32 | a << 16 | b, reasons << 32 | k0}.
GRR_DEV void flag_push(unsigned long long *flist, unsigned long long fcap, long long e, int a, int b, int k0, int reasons) {
#ifdef __CUDA_ARCH__
  const unsigned long long pos = atomicAdd(flist, 1ull);
  if (pos < fcap) {
    flist[2 + 2 * pos] = ((unsigned long long)e << 32) | ((unsigned long long)(unsigned)a << 16) | (unsigned long long)(unsigned)b;
    flist[3 + 2 * pos] = ((unsigned long long)(unsigned)reasons << 32) | (unsigned long long)(unsigned)k0;
  }
#endif
}

// explicit pairs, SoA qa[n][P], qb[n][P]  (validEdge, graph.py:1051-1053), work-fetching lanes
template <typename Real, int M, bool PL, bool SPIN = false>
__global__ void __launch_bounds__(256)
k_valid_edge_batch(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long total, const Real *__restrict__ qa,
                   const Real *__restrict__ qb, const Real *__restrict__ wa, const Real *__restrict__ wb, Real epsilon,
                   uint8_t *__restrict__ valid, int32_t *__restrict__ info, unsigned long long *work,
                   unsigned long long *counters) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(reinterpret_cast<Real *>(smem_raw) + threadIdx.x, (int)blockDim.x, n);
  const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));          // graph.py:945
  const Real thr = eps * Real(c.nlinks_eps);                     // graph.py:947
  const Real thr2 = thr * thr;
  LaneIK<Real, M> L;
  lane_begin_mem(L, mem);
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  PairLane<Real> P;
  WInterp<Real, M> W;
  Target<Real, M> tg;
  tg.x[0] = tg.x[1] = tg.x[2] = 0;
  if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
  LaneCounters cnt;
  long long t = -1;
  bool done = false;
  for (;;) {
    int verdict = -1, fail = 0;
    if (t < 0 && !done) {
      const long long nt = (long long)atomicAdd(work, 1ull);
      if (nt >= total) done = true;
      else {
        t = nt;
        Real a6[6], b6[6];
        for (int k = 0; k < dw; k++) { a6[k] = wa[t * dw + k]; b6[k] = wb[t * dw + k]; }
        winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
        P.qa = qa + t; P.qb = qb + t; P.sa = (int)total; P.sb = (int)total;
        if (!pair_begin<Real, M>(c, eps, P, L, kspin)) verdict = 1;
        else winterp_at<Real, M>(W, P.u_at(1), tg);
      }
    }
    const bool running = (t >= 0 && verdict < 0);
    if (__all_sync(0xffffffffu, t < 0)) break;
    if (__any_sync(0xffffffffu, running)) {
      PassOut<Real, M> o;
      const bool first = running && P.k == 1;
      if constexpr (PL) lane_pass_planar<Real, true>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      else lane_pass<Real, M, true, SPIN>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      const int st = lane_step<Real, M, PL>(c, L, mem, o, running);
      if (running && st != IK_RUNNING) {
        P.solves++; P.iters += L.it;
        cnt.solves++; cnt.iters += L.it; cnt.evals += L.evals; cnt.ok += (st == IK_OK);
        if (st != IK_OK) { verdict = 0; fail = 1; }               // graph.py:964-966
        else {
          verdict = pair_advance(c, mem, o.s2, thr2, P, L, false, kspin);
          if (verdict == 0) fail = 2;
          if (verdict < 0) winterp_at<Real, M>(W, P.u_at(P.k), tg);
        }
      }
    }
    if (t >= 0 && verdict >= 0) {
      valid[t] = (uint8_t)verdict;
      if (info) { info[4 * t] = P.ndivs; info[4 * t + 1] = P.solves; info[4 * t + 2] = fail; info[4 * t + 3] = P.iters; }
      cnt.pairs++; cnt.valid += verdict;
      t = -1;
      lane_begin_mem(L, mem);     // idle lanes keep walking valid memory
    }
  }
  cnt.flush(counters);
}

// ------------------------------------------------------------------------------------------
// E1/E3: all-pairs edge construction, one CTA per workspace edge
// (connectConfigurationSpaceOnEdge solver.py:689-697 under connectConfigurationSpace :700-755).
// Both endpoint nodes' config blocks are staged in shared memory; lanes pull pair indices from a
// shared-memory counter and run their pairs as independent lane machines; verdict bits are OR-ed into
// a shared-memory Nq x Nq bitmask that is written out once.
// ------------------------------------------------------------------------------------------
// KT > 0: instantiation for CTAs of exactly KT lanes (planar chains; see lane_pass_planar_)
template <typename Real, int M, bool PL, bool FLAG, bool SPIN = false, int KT = 0>
#ifndef GRR_KT_MIN_BLOCKS
#define GRR_KT_MIN_BLOCKS 2
#endif
// (the near-tie build of the KT instantiation may use the registers two CTAs of KT lanes leave it -- 153 instead of 119:
// pass loop 140 -> 134 instructions, config 2 exact build 0.4086 -> 0.4058 s; the plain build is 2 % slower that way and keeps
// the default bound)
#if GRR_KT_MIN_BLOCKS
__global__ void __launch_bounds__(KT > 0 && FLAG ? KT : (sizeof(Real) == 4 ? (M == 3 ? 512 : 384) : 256), KT > 0 && FLAG ? GRR_KT_MIN_BLOCKS : 0)
#else
__global__ void __launch_bounds__(sizeof(Real) == 4 ? (M == 3 ? 512 : 384) : 256)
#endif
k_edges_build(const __grid_constant__ ChainDev<Real, kMaxJ> c, const int32_t *__restrict__ wedges,
              const Real *__restrict__ params, const Real *__restrict__ q_kept, const int32_t *__restrict__ count,
              const int32_t *__restrict__ old_count, int Nq, int Nqp, Real epsilon, uint32_t *mask,
              uint32_t *__restrict__ edge_stats, unsigned long long *counters, unsigned long long *flist,
              unsigned long long fcap) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  const int Wd = (Nq + 31) >> 5;
  Real *sa = reinterpret_cast<Real *>(smem_raw);
  Real *sb = sa + n * Nqp;
  uint32_t *smask = reinterpret_cast<uint32_t *>(sb + n * Nqp);
  Real *lanes = reinterpret_cast<Real *>(smask + ((Nq * Wd + 3) & ~3));
  __shared__ unsigned s_next, s_tested, s_valid, s_bits, s_ndsum, s_ndcnt;
  __shared__ uint2 s_queue[16 * 32];          // per-warp buffers of claimed pairs (<= 512 lanes per CTA)
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
  const int ca = count[iw], cb = count[jw];
  const int oa = old_count ? old_count[iw] : 0, ob = old_count ? old_count[jw] : 0;
  uint32_t *mrow = mask + e * (long long)Nq * Wd;
  const unsigned total = (unsigned)ca * (unsigned)cb;
  if (threadIdx.x == 0) { s_next = 0; s_tested = 0; s_valid = 0; s_bits = 0; s_ndsum = 0; s_ndcnt = 0; }
  // mask in shared memory: old verdicts of both-old pairs are kept (incremental growth, solver.py:694)
  for (int k = threadIdx.x; k < Nq * Wd; k += blockDim.x) {
    uint32_t w = 0;
    if (old_count) {
      const int a = k / Wd, b0 = (k - a * Wd) * 32;
      if (a < oa && b0 < ob) { w = mrow[k]; if (ob - b0 < 32) w &= (1u << (ob - b0)) - 1u; }
    }
    smask[k] = w;
  }
  if (total) {
    const Real *ga = q_kept + (long long)iw * n * Nqp, *gb = q_kept + (long long)jw * n * Nqp;
    for (int k = threadIdx.x; k < n * Nqp; k += blockDim.x) { sa[k] = ga[k]; sb[k] = gb[k]; }
  }
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(lanes + threadIdx.x, (int)blockDim.x, n);
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  __syncthreads();
  LaneCounters cnt;
  if (total) {
    const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));
    const Real thr = eps * Real(c.nlinks_eps);
    const Real thr2 = thr * thr;
    WInterp<Real, M> W;
    {
      Real a6[6], b6[6];
      for (int k = 0; k < dw; k++) { a6[k] = params[(long long)iw * dw + k]; b6[k] = params[(long long)jw * dw + k]; }
      winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
    }
    LaneIK<Real, M> L;
    lane_begin_mem(L, mem);
    PairLane<Real> P;
    Target<Real, M> tg;
    tg.x[0] = tg.x[1] = tg.x[2] = 0;
    if (M == 6) { for (int k = 0; k < 9; k++) tg.R[k] = (k % 4 == 0) ? Real(1) : Real(0); }
    // Longest pairs first: the edge is swept twice, first for the pairs whose Ndivs is at least the mean of a sample
    // (taken here, one pair per thread), then for the shorter ones, so that the pairs still running when the edge
    // runs dry are short ones (6 % of all warp time was spent waiting at the end-of-edge barrier in v8).
    {
      const unsigned ts = (unsigned)(((unsigned long long)threadIdx.x * total) / blockDim.x);
      const int ta = (int)(ts / (unsigned)cb), tb = (int)(ts - (unsigned)ta * (unsigned)cb);
      const int nd = pair_ndivs<Real>(n, kspin, eps, sa + ta, Nqp, sb + tb, Nqp);
      const unsigned sum = __reduce_add_sync(0xffffffffu, (unsigned)max(nd, 0));
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_ndsum, sum); atomicAdd(&s_ndcnt, 32u); }
    }
    __syncthreads();
    const int nd_cut = (int)((s_ndsum + s_ndcnt / 2) / s_ndcnt);
    const unsigned total2 = 2u * total;
    int a = -1, b = 0;
    // Pairs are claimed 32 at a time per warp: all lanes compute the joint-space distance (-> Ndivs) of one
    // claimed pair each, in lock step, and the pairs that need interpolated solves go to a per-warp buffer from
    // which lanes take one whenever they become free.  (A lane claiming and measuring its own pair did so with
    // 1.5 of 32 lanes active: 5 % of all issued instructions in v6.)
    const int lane = threadIdx.x & 31;
    uint2 *queue = reinterpret_cast<uint2 *>(s_queue) + (threadIdx.x >> 5) * 32;
    int qhead = 0, qtail = 0;          // warp-uniform
    bool exhausted = false;            // warp-uniform: the edge has no unclaimed pairs left
    for (;;) {
      int verdict = -1;
      unsigned needm = __ballot_sync(0xffffffffu, a < 0);
      while (needm) {
        if (qhead == qtail) {
          if (exhausted) break;
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(&s_next, 32u);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (base >= total2) { exhausted = true; break; }
          const unsigned t2 = base + lane;
          const bool second = t2 >= total;                        // second sweep: the shorter pairs
          const unsigned t = second ? t2 - total : t2;
          bool ok = t2 < total2;
          int ta = 0, tb = 0, nd = 0;
          if (ok) {
            ta = (int)(t / (unsigned)cb); tb = (int)(t - (unsigned)ta * (unsigned)cb);
            if (ta < oa && tb < ob) ok = false;                   // solver.py:694
            if (iw == jw && ta <= tb) ok = false;                 // self-motion pairs: validEdge(q_i, q_j, x, x), i > j (solver.py:648-653)
          }
          if (ok) {
            nd = pair_ndivs<Real, FLAG>(n, kspin, eps, sa + ta, Nqp, sb + tb, Nqp, c.fl_nd);
            if ((nd & ~kNdFlag) <= 0) {                           // dist == 0 -> valid, no solves
              if (!second) {
                atomicOr(&smask[ta * Wd + (tb >> 5)], 1u << (tb & 31)); cnt.pairs++; cnt.valid++;
                if (FLAG && (nd & kNdFlag)) flag_push(flist, fcap, e, ta, tb, 0, PH_R_NDIVS);
              }
              ok = false;
            } else if (((nd & ~kNdFlag) >= nd_cut) == second) ok = false;      // not this sweep's
          }
          const unsigned okm = __ballot_sync(0xffffffffu, ok);
          if (ok) queue[__popc(okm & ((1u << lane) - 1u))] = make_uint2(((unsigned)ta << 16) | (unsigned)tb, (unsigned)nd);
          __syncwarp();
          qhead = 0; qtail = __popc(okm);
          continue;
        }
        const int rank = __popc(needm & ((1u << lane) - 1u)), avail = qtail - qhead;
        if (a < 0 && rank < avail) {
          const uint2 ent = queue[qhead + rank];
          a = (int)(ent.x >> 16); b = (int)(ent.x & 0xffffu);
          P.qa = sa + a; P.qb = sb + b; P.sa = Nqp; P.sb = Nqp;
          pair_arm<Real, M>((int)ent.y, P, L);
          winterp_at<Real, M>(W, P.u_at(1), tg);
        }
        __syncwarp();
        qhead += min(__popc(needm), avail);
        needm = __ballot_sync(0xffffffffu, a < 0);
      }
      const bool running = (a >= 0 && verdict < 0);
      if (__all_sync(0xffffffffu, a < 0)) break;
      if (__any_sync(0xffffffffu, running)) {
        PassOut<Real, M> o;
        const bool first = running && P.k == 1;
        if constexpr (PL) lane_pass_planar<Real, true, KT>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
        else lane_pass<Real, M, true, SPIN>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
        const int st = lane_step<Real, M, PL, FLAG>(c, L, mem, o, running);
        if constexpr (FLAG) warp_tie_service<Real, M, PL>(c, L, mem, running);
        constexpr bool COOP = sizeof(Real) == 4;                // (double precision keeps the sequential sum of the oracle)
        if (running && st != IK_RUNNING) {
          cnt.solves++; cnt.iters += L.it; cnt.evals += L.evals; cnt.ok += (st == IK_OK);
          if (FLAG) pair_note_solve(P, L, st);
          if (st != IK_OK) verdict = 0;                         // graph.py:964-966
          else {
            verdict = pair_advance<FLAG, COOP>(c, mem, o.s2, thr2, P, L, false, kspin);
            if (verdict == -1) winterp_at<Real, M>(W, P.u_at(P.k), tg);
          }
        }
        if constexpr (COOP) {
          const Real e2 = warp_close_service<Real, M>(n, kspin, mem.T, P, L, running && verdict == kNeedClose);
          if (running && verdict == kNeedClose) verdict = pair_close<FLAG>(c, e2, thr2, P, L);
        }
      }
      if (a >= 0 && verdict >= 0) {
        if (verdict) atomicOr(&smask[a * Wd + (b >> 5)], 1u << (b & 31));
        if (FLAG && P.flag) flag_push(flist, fcap, e, a, b, P.k0, P.flag);
        cnt.pairs++; cnt.valid += verdict;
        a = -1;
        lane_begin_mem(L, mem);   // idle lanes keep walking valid memory
      }
    }
  }
  {
    const unsigned tp = __reduce_add_sync(0xffffffffu, cnt.pairs), tv = __reduce_add_sync(0xffffffffu, cnt.valid);
    if ((threadIdx.x & 31) == 0 && tp) { atomicAdd(&s_tested, tp); atomicAdd(&s_valid, tv); }
  }
  cnt.flush(counters);
  __syncthreads();
  unsigned nbits = 0;
  for (int k = threadIdx.x; k < Nq * Wd; k += blockDim.x) { const uint32_t w = smask[k]; mrow[k] = w; nbits += __popc(w); }
  if (edge_stats) {
    // {pairs tested in this call, valid bits now stored for the edge}
    const unsigned wbits = __reduce_add_sync(0xffffffffu, nbits);
    if ((threadIdx.x & 31) == 0 && wbits) atomicAdd(&s_bits, wbits);
    __syncthreads();
    if (threadIdx.x == 0) { edge_stats[2 * e] = s_tested; edge_stats[2 * e + 1] = s_bits; }
  }
}

// ------------------------------------------------------------------------------------------
// E1/E3, dense roadmaps of the planar small-angle chains (BASELINE config 2): k_edges_build with TWO lane machines
// (slots) per thread on the packed FP32 instructions (grr_lane2.cuh).  Same claiming, queueing, verdict and near-tie
// logic per slot; the chain pass serves both slots at once.
// ------------------------------------------------------------------------------------------
template <bool FLAG>
__global__ void __launch_bounds__(256)
k_edges_build2(const __grid_constant__ ChainDev<float, kMaxJ> c, const int32_t *__restrict__ wedges,
               const float *__restrict__ params, const float *__restrict__ q_kept, const int32_t *__restrict__ count,
               const int32_t *__restrict__ old_count, int Nq, int Nqp, float epsilon, uint32_t *mask,
               uint32_t *__restrict__ edge_stats, unsigned long long *counters, unsigned long long *flist,
               unsigned long long fcap) {
  using Real = float;
  constexpr int M = 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n;
  const int Wd = (Nq + 31) >> 5;
  Real *sa = reinterpret_cast<Real *>(smem_raw);
  Real *sb = sa + n * Nqp;
  uint32_t *smask = reinterpret_cast<uint32_t *>(sb + n * Nqp);
  Real *lanes = reinterpret_cast<Real *>(smask + ((Nq * Wd + 3) & ~3));
  __shared__ unsigned s_next, s_tested, s_valid, s_bits, s_ndsum, s_ndcnt;
  __shared__ uint2 s_queue[8 * 32];           // per-warp buffers of claimed pairs (<= 256 threads per CTA)
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
  const int ca = count[iw], cb = count[jw];
  const int oa = old_count ? old_count[iw] : 0, ob = old_count ? old_count[jw] : 0;
  uint32_t *mrow = mask + e * (long long)Nq * Wd;
  const unsigned total = (unsigned)ca * (unsigned)cb;
  if (threadIdx.x == 0) { s_next = 0; s_tested = 0; s_valid = 0; s_bits = 0; s_ndsum = 0; s_ndcnt = 0; }
  for (int k = threadIdx.x; k < Nq * Wd; k += blockDim.x) {
    uint32_t w = 0;
    if (old_count) {
      const int a = k / Wd, b0 = (k - a * Wd) * 32;
      if (a < oa && b0 < ob) { w = mrow[k]; if (ob - b0 < 32) w &= (1u << (ob - b0)) - 1u; }
    }
    smask[k] = w;
  }
  if (total) {
    const Real *ga = q_kept + (long long)iw * n * Nqp, *gb = q_kept + (long long)jw * n * Nqp;
    for (int k = threadIdx.x; k < n * Nqp; k += blockDim.x) { sa[k] = ga[k]; sb[k] = gb[k]; }
  }
  LaneMem<Real, 2> mem[2];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    mem[s].init(lanes + 2 * threadIdx.x + s, 2 * (int)blockDim.x, n);
    for (int j = 0; j < lane_reals<2>(n); j++) mem[s].base[j * mem[s].T] = 0;
  }
  __syncthreads();
  LaneCounters cnt;
  if (total) {
    const Real eps = epsilon * sqrt_(Real(c.nlinks_eps));
    const Real thr = eps * Real(c.nlinks_eps);
    const Real thr2 = thr * thr;
    WInterp<Real, M> W;
    {
      Real a6[6], b6[6];
      for (int k = 0; k < 3; k++) { a6[k] = params[(long long)iw * 3 + k]; b6[k] = params[(long long)jw * 3 + k]; }
      winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
    }
    LaneIK<Real, M> L[2];
    PairLane<Real> P[2];
    Target<Real, M> tg[2];
    int a[2] = {-1, -1}, b[2] = {0, 0};
#pragma unroll
    for (int s = 0; s < 2; s++) { lane_begin_mem(L[s], mem[s]); tg[s].x[0] = tg[s].x[1] = tg[s].x[2] = 0; P[s].qa = sa; P[s].qb = sb; P[s].sa = Nqp; P[s].sb = Nqp; P[s].k = 2; }
    {
      const unsigned ts = (unsigned)(((unsigned long long)threadIdx.x * total) / blockDim.x);
      const int ta = (int)(ts / (unsigned)cb), tb = (int)(ts - (unsigned)ta * (unsigned)cb);
      const int nd = pair_ndivs<Real>(n, 0u, eps, sa + ta, Nqp, sb + tb, Nqp);
      const unsigned sum = __reduce_add_sync(0xffffffffu, (unsigned)max(nd, 0));
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_ndsum, sum); atomicAdd(&s_ndcnt, 32u); }
    }
    __syncthreads();
    const int nd_cut = (int)((s_ndsum + s_ndcnt / 2) / s_ndcnt);
    const unsigned total2 = 2u * total;
    const int lane = threadIdx.x & 31;
    uint2 *queue = reinterpret_cast<uint2 *>(s_queue) + (threadIdx.x >> 5) * 32;
    int qhead = 0, qtail = 0;          // warp-uniform
    bool exhausted = false;            // warp-uniform
    for (;;) {
      int verdict[2] = {-1, -1};
#pragma unroll
      for (int s = 0; s < 2; s++) {
        unsigned needm = __ballot_sync(0xffffffffu, a[s] < 0);
        while (needm) {
          if (qhead == qtail) {
            if (exhausted) break;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&s_next, 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= total2) { exhausted = true; break; }
            const unsigned t2 = base + lane;
            const bool second = t2 >= total;
            const unsigned t = second ? t2 - total : t2;
            bool ok = t2 < total2;
            int ta = 0, tb = 0, nd = 0;
            if (ok) {
              ta = (int)(t / (unsigned)cb); tb = (int)(t - (unsigned)ta * (unsigned)cb);
              if (ta < oa && tb < ob) ok = false;                   // solver.py:694
              if (iw == jw && ta <= tb) ok = false;                 // self-motion pairs i > j (solver.py:648-653)
            }
            if (ok) {
              nd = pair_ndivs<Real, FLAG>(n, 0u, eps, sa + ta, Nqp, sb + tb, Nqp, c.fl_nd);
              if ((nd & ~kNdFlag) <= 0) {
                if (!second) {
                  atomicOr(&smask[ta * Wd + (tb >> 5)], 1u << (tb & 31)); cnt.pairs++; cnt.valid++;
                  if (FLAG && (nd & kNdFlag)) flag_push(flist, fcap, e, ta, tb, 0, PH_R_NDIVS);
                }
                ok = false;
              } else if (((nd & ~kNdFlag) >= nd_cut) == second) ok = false;
            }
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            if (ok) queue[__popc(okm & ((1u << lane) - 1u))] = make_uint2(((unsigned)ta << 16) | (unsigned)tb, (unsigned)nd);
            __syncwarp();
            qhead = 0; qtail = __popc(okm);
            continue;
          }
          const int rank = __popc(needm & ((1u << lane) - 1u)), avail = qtail - qhead;
          if (a[s] < 0 && rank < avail) {
            const uint2 ent = queue[qhead + rank];
            a[s] = (int)(ent.x >> 16); b[s] = (int)(ent.x & 0xffffu);
            P[s].qa = sa + a[s]; P[s].qb = sb + b[s]; P[s].sa = Nqp; P[s].sb = Nqp;
            pair_arm<Real, M>((int)ent.y, P[s], L[s]);
            winterp_at<Real, M>(W, P[s].u_at(1), tg[s]);
          }
          __syncwarp();
          qhead += min(__popc(needm), avail);
          needm = __ballot_sync(0xffffffffu, a[s] < 0);
        }
      }
      if (__all_sync(0xffffffffu, a[0] < 0 && a[1] < 0)) break;
      const bool run0 = a[0] >= 0, run1 = a[1] >= 0;
      {
        PassOut<Real, M> o[2];
        const bool f0 = run0 && P[0].k == 1, f1 = run1 && P[1].k == 1;
        const float *const xref[2] = {f0 ? P[0].qa : mem[0].prev(), f1 ? P[1].qa : mem[1].prev()};
        const int sref[2] = {f0 ? P[0].sa : mem[0].T, f1 ? P[1].sa : mem[1].T};
        lane_pass_planar2(c, tg, L, mem, xref, sref, o);
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const bool running = (s == 0 ? run0 : run1);
          const int st = lane_step<Real, M, true, FLAG>(c, L[s], mem[s], o[s], running);
          if constexpr (FLAG) warp_tie_service<Real, M, true, 2>(c, L[s], mem[s], running);
          if (running && st != IK_RUNNING) {
            cnt.solves++; cnt.iters += L[s].it; cnt.evals += L[s].evals; cnt.ok += (st == IK_OK);
            if (FLAG) pair_note_solve(P[s], L[s], st);
            if (st != IK_OK) verdict[s] = 0;                       // graph.py:964-966
            else {
              verdict[s] = pair_advance<FLAG>(c, mem[s], o[s].s2, thr2, P[s], L[s]);
              if (verdict[s] < 0) winterp_at<Real, M>(W, P[s].u_at(P[s].k), tg[s]);
            }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (a[s] >= 0 && verdict[s] >= 0) {
          if (verdict[s]) atomicOr(&smask[a[s] * Wd + (b[s] >> 5)], 1u << (b[s] & 31));
          if (FLAG && P[s].flag) flag_push(flist, fcap, e, a[s], b[s], P[s].k0, P[s].flag);
          cnt.pairs++; cnt.valid += verdict[s];
          a[s] = -1;
          lane_begin_mem(L[s], mem[s]);   // idle slots keep walking valid memory
          P[s].k = 2;                     // (their reference config is the slot's own prev buffer)
        }
      }
    }
  }
  {
    const unsigned tp = __reduce_add_sync(0xffffffffu, cnt.pairs), tv = __reduce_add_sync(0xffffffffu, cnt.valid);
    if ((threadIdx.x & 31) == 0 && tp) { atomicAdd(&s_tested, tp); atomicAdd(&s_valid, tv); }
  }
  cnt.flush(counters);
  __syncthreads();
  unsigned nbits = 0;
  for (int k = threadIdx.x; k < Nq * Wd; k += blockDim.x) { const uint32_t w = smask[k]; mrow[k] = w; nbits += __popc(w); }
  if (edge_stats) {
    const unsigned wbits = __reduce_add_sync(0xffffffffu, nbits);
    if ((threadIdx.x & 31) == 0 && wbits) atomicAdd(&s_bits, wbits);
    __syncthreads();
    if (threadIdx.x == 0) { edge_stats[2 * e] = s_tested; edge_stats[2 * e + 1] = s_bits; }
  }
}

// ------------------------------------------------------------------------------------------
// E1/E3 for SPARSE roadmaps (a few configs per node: 6-D tasks, arms with narrow reach -- BASELINE configs
// 3-5 have 18-200 pair tests per workspace edge).  A CTA per workspace edge then holds 1-3 pairs per lane and
// most lanes idle while the longest pair finishes.  Here the pairs of ALL workspace edges form one queue:
// offs[e] = pairs of the edges before e (exclusive scan of count[iw]*count[jw]); persistent lanes run them as
// independent lane machines.  A warp takes 32 consecutive queue positions at a time, decodes them together
// (binary search in offs, all lanes at once) into a per-warp shared-memory buffer, and hands them to its lanes
// as they become free, so the decode costs ~1/32 of a search per pair and no lane waits for another lane's
// pair.  Endpoint configs are read in place from Q (the few configs of a node stay in L1); verdict bits are
// OR-ed into the global mask, which k_edges_prepare cleared (keeping both-old bits, solver.py:694).
// ------------------------------------------------------------------------------------------
static __global__ void k_edges_prepare(long long E, const int32_t *__restrict__ wedges, const int32_t *__restrict__ count,
                                       const int32_t *__restrict__ old_count, int Nq, int Wd, uint32_t *mask,
                                       unsigned long long *offs, uint32_t *edge_stats) {
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
  const int ca = count[iw], cb = count[jw];
  const int oa = old_count ? old_count[iw] : 0, ob = old_count ? old_count[jw] : 0;
  uint32_t *mrow = mask + e * (long long)Nq * Wd;
  for (int k = threadIdx.x; k < Nq * Wd; k += blockDim.x) {
    uint32_t w = 0;
    if (old_count) {
      const int a = k / Wd, b0 = (k - a * Wd) * 32;
      if (a < oa && b0 < ob) { w = mrow[k]; if (ob - b0 < 32) w &= (1u << (ob - b0)) - 1u; }
    }
    mrow[k] = w;
  }
  if (threadIdx.x == 0) {
    offs[e + 1] = (unsigned long long)ca * (unsigned long long)cb;     // queue positions (both-old ones are skipped on decode)
    if (e == 0) offs[0] = 0;
    if (edge_stats) edge_stats[2 * e] = (iw == jw) ? (uint32_t)(((long long)ca * (ca - 1) - (long long)oa * (oa - 1)) / 2)
                                                   : (uint32_t)((long long)ca * cb - (long long)oa * ob);
  }
}

// in-place inclusive scan of a[1..N] (a[0] = 0 stays): one CTA of 1024 threads, a contiguous chunk per thread
static __global__ void __launch_bounds__(1024) k_scan_u64(long long N, unsigned long long *a) {
  __shared__ unsigned long long part[1024];
  const long long chunk = (N + 1023) / 1024;
  const long long lo = 1 + (long long)threadIdx.x * chunk, hi = min(lo + chunk, N + 1);
  unsigned long long s = 0;
  for (long long k = lo; k < hi; k++) s += a[k];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned long long v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = part[threadIdx.x] - s;
  for (long long k = lo; k < hi; k++) { run += a[k]; a[k] = run; }
}

// valid bits now stored per workspace edge -> edge_stats[e][1]
static __global__ void k_edges_bits(long long E, int words, const uint32_t *__restrict__ mask, uint32_t *edge_stats) {
  const long long e = blockIdx.x;
  const uint32_t *mrow = mask + e * (long long)words;
  unsigned nb = 0;
  for (int k = threadIdx.x; k < words; k += blockDim.x) nb += __popc(mrow[k]);
  nb = __reduce_add_sync(0xffffffffu, nb);
  __shared__ unsigned s_bits;
  if (threadIdx.x == 0) s_bits = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && nb) atomicAdd(&s_bits, nb);
  __syncthreads();
  if (threadIdx.x == 0) edge_stats[2 * e + 1] = s_bits;
}

// double precision: 3 CTAs per SM for 3 residual rows (168 registers, 36 bytes of spills: 12 instead of 8 warps per SM hide
// the FP64 latencies -- the kernel sat at 52 % issue utilisation with `stalled_wait` 1.6 per issue), 2 for 6 rows
// (capping that build at 168 registers spills 1.6 KB per thread)
template <typename Real, int M, bool PL, bool FLAG, bool SPIN = false>
// the queue kernel always runs CTAs of 128 lanes: the lane stride of its passes is a compile-time constant (0 = run-time)
#ifndef GRR_FLAT_KT
#define GRR_FLAT_KT 128
#endif
#ifndef GRR_F64M3_BLOCKS
#define GRR_F64M3_BLOCKS 3
#endif
#ifndef GRR_F64M6_BLOCKS
#define GRR_F64M6_BLOCKS 2
#endif
__global__ void __launch_bounds__(128, (sizeof(Real) == 4 ? (M == 3 ? 4 : 3) : (M == 3 ? GRR_F64M3_BLOCKS : GRR_F64M6_BLOCKS)))
k_edges_flat(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long E, const int32_t *__restrict__ wedges,
             const Real *__restrict__ params, const Real *__restrict__ q_kept, const int32_t *__restrict__ count,
             const int32_t *__restrict__ old_count, int Nq, int Nqp, Real epsilon, uint32_t *mask,
             const unsigned long long *__restrict__ offs, unsigned long long *work, unsigned long long *counters,
             unsigned long long *flist, unsigned long long fcap) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  const int Wd = (Nq + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per-warp buffer of decoded queue positions: {workspace edge, a << 16 | b, Ndivs}
  int4 *queue = reinterpret_cast<int4 *>(smem_raw) + warp * 32;
  Real *lanes = reinterpret_cast<Real *>(reinterpret_cast<int4 *>(smem_raw) + (blockDim.x >> 5) * 32);
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(lanes + threadIdx.x, (int)blockDim.x, n);
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  const unsigned long long total = offs[E];
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
  LaneCounters cnt;
  long long edge = -1;               // this lane's pair: workspace edge, (a, b)
  int a = 0, b = 0;
  int qhead = 0, qtail = 0;          // warp-uniform
  bool exhausted = false;            // warp-uniform: the global queue has no positions left
  __syncwarp();
  for (;;) {
    int verdict = -1;
    // ---- hand queue positions to the idle lanes of this warp ----
    unsigned needm = __ballot_sync(0xffffffffu, edge < 0);
    while (needm) {
      if (qhead == qtail) {
        if (exhausted) break;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(work, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) { exhausted = true; break; }
        const unsigned long long t = base + lane;
        bool ok = t < total;
        long long e = 0;
        int ta = 0, tb = 0, nd = 0;
        if (ok) {
          long long lo = 0, hi = E;                              // offs[lo] <= t < offs[hi]
          while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (offs[mid] <= t) lo = mid; else hi = mid; }
          e = lo;
          const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
          const unsigned cb = (unsigned)count[jw];
          const unsigned off = (unsigned)(t - offs[e]);
          ta = (int)(off / cb); tb = (int)(off - (unsigned)ta * cb);
          GRR_DCHECK(e >= 0 && e < E && offs[e] <= t && t < offs[e + 1] && cb > 0 && ta < count[iw] && tb < (int)cb && ta < Nq && tb < Nq);
          if (old_count && ta < old_count[iw] && tb < old_count[jw]) ok = false;      // solver.py:694
          if (iw == jw && ta <= tb) ok = false;                                       // self-motion pairs i > j (solver.py:648-653)
          if (ok) {
            // all lanes measure their pair together (a lane doing this alone ran with 1.5 of 32 lanes active)
            nd = pair_ndivs<Real, FLAG>(n, kspin, eps, q_kept + (long long)iw * n * Nqp + ta, Nqp, q_kept + (long long)jw * n * Nqp + tb, Nqp, c.fl_nd);
            if ((nd & ~kNdFlag) <= 0) {                           // dist == 0 -> valid, no solves
              atomicOr(&mask[(e * Nq + ta) * Wd + (tb >> 5)], 1u << (tb & 31));
              cnt.pairs++; cnt.valid++;
              if (FLAG && (nd & kNdFlag)) flag_push(flist, fcap, e, ta, tb, 0, PH_R_NDIVS);
              ok = false;
            }
          }
        }
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (ok) queue[__popc(okm & ((1u << lane) - 1u))] = make_int4((int)e, (int)(((unsigned)ta << 16) | (unsigned)tb), nd, 0);
        __syncwarp();
        qhead = 0; qtail = __popc(okm);
        continue;
      }
      const int rank = __popc(needm & ((1u << lane) - 1u)), avail = qtail - qhead;
      if (edge < 0 && rank < avail) {
        GRR_DCHECK(qhead + rank < 32 && qtail <= 32);
        const int4 ent = queue[qhead + rank];
        edge = ent.x; a = (int)((unsigned)ent.y >> 16); b = ent.y & 0xffff;
        const int iw = wedges[2 * edge], jw = wedges[2 * edge + 1];
        Real a6[6], b6[6];
        for (int k = 0; k < dw; k++) { a6[k] = params[(long long)iw * dw + k]; b6[k] = params[(long long)jw * dw + k]; }
        winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
        P.qa = q_kept + (long long)iw * n * Nqp + a; P.qb = q_kept + (long long)jw * n * Nqp + b; P.sa = Nqp; P.sb = Nqp;
        pair_arm<Real, M>(ent.z, P, L);
        winterp_at<Real, M>(W, P.u_at(1), tg);
      }
      __syncwarp();
      qhead += min(__popc(needm), avail);
      needm = __ballot_sync(0xffffffffu, edge < 0);
    }
    if (__all_sync(0xffffffffu, edge < 0)) break;                // nothing running and nothing left to hand out
    const bool running = (edge >= 0 && verdict < 0);
    if (__any_sync(0xffffffffu, running)) {
      PassOut<Real, M> o;
      const bool first = running && P.k == 1;
      if constexpr (PL) lane_pass_planar<Real, true, GRR_FLAT_KT>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      else lane_pass<Real, M, true, SPIN, GRR_FLAT_KT>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      const int st = lane_step<Real, M, PL, FLAG>(c, L, mem, o, running);
      if constexpr (FLAG) warp_tie_service<Real, M, PL>(c, L, mem, running);
      if (running && st != IK_RUNNING) {
        cnt.solves++; cnt.iters += L.it; cnt.evals += L.evals; cnt.ok += (st == IK_OK);
        if (FLAG) pair_note_solve(P, L, st);
        if (st != IK_OK) verdict = 0;                             // graph.py:964-966
        else {
          verdict = pair_advance<FLAG>(c, mem, o.s2, thr2, P, L, false, kspin);
          if (verdict < 0) winterp_at<Real, M>(W, P.u_at(P.k), tg);
        }
      }
    }
    if (edge >= 0 && verdict >= 0) {
      GRR_DCHECK(edge < E && a < Nq && b < Nq);
      if (verdict) atomicOr(&mask[(edge * Nq + a) * Wd + (b >> 5)], 1u << (b & 31));
      if (FLAG && P.flag) flag_push(flist, fcap, edge, a, b, P.k0, P.flag);
      cnt.pairs++; cnt.valid += verdict;
      edge = -1;
      lane_begin_mem(L, mem);   // idle lanes keep walking valid memory
    }
  }
  cnt.flush(counters);
}

// ------------------------------------------------------------------------------------------
// Second pass of the exact edge build (grr_edges_build_exact): the pairs the single-precision pass listed as
// near-ties are decided again in the arithmetic of the reference (IEEE double, on the double-precision configs)
// and their mask bits overwritten.  Solves 1 .. k0 - 1 of a listed pair ran clear of every margin, so the
// double-precision run restarts at solve max(k0 - 1, 1) (k0 - 1 supplies the predecessor of the leaf (k0 - 1, k0);
// the leaf before it was settled by the first pass) and runs the pair to its end.  k0 = 0 (Ndivs rounding): whole pair.
// Persistent lanes pull list entries from a global counter.
// ------------------------------------------------------------------------------------------
template <typename Real, int M, bool PL, bool SPIN = false>
__global__ void __launch_bounds__(256)
k_edges_recheck(const __grid_constant__ ChainDev<Real, kMaxJ> c, const int32_t *__restrict__ wedges,
                const Real *__restrict__ params, const Real *__restrict__ q_kept, int Nq, int Nqp, Real epsilon, uint32_t *mask,
                uint32_t *edge_stats, const unsigned long long *__restrict__ flist, unsigned long long fcap,
                unsigned long long *work, unsigned long long *counters, int listed_mode) {
  const unsigned kspin = SPIN ? c.spin_mask : 0u;      // compile-time 0 in the kernels of chains without spin joints
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = c.n, dw = (c.mode == 2 ? 6 : 3);
  const int Wd = (Nq + 31) >> 5;
  LaneMem<Real, col_rows(M, PL)> mem;
  mem.init(reinterpret_cast<Real *>(smem_raw) + threadIdx.x, (int)blockDim.x, n);
  for (int j = 0; j < lane_reals<col_rows(M, PL)>(n); j++) mem.base[j * mem.T] = 0;
  LaneCounters cnt;
  const unsigned long long listed = flist[0];
  const unsigned long long total = listed < fcap ? listed : fcap;
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
  long long edge = -1;
  int a = 0, b = 0, kstart = 1;
  unsigned flagged = 0, flipped = 0;
  long long dvalid = 0;
  bool done = false;
  for (;;) {
    int verdict = -1;
    if (edge < 0 && !done) {
      const unsigned long long t = atomicAdd(work, 1ull);
      if (t >= total) done = true;
      else {
        const unsigned long long w0 = flist[2 + 2 * t], w1 = flist[3 + 2 * t];
        edge = (long long)(w0 >> 32); a = (int)((w0 >> 16) & 0xffffu); b = (int)(w0 & 0xffffu);
        GRR_DCHECK(a < Nq && b < Nq && edge >= 0);
        const int k0 = (int)(w1 & 0xffffffffu);
        const int iw = wedges[2 * edge], jw = wedges[2 * edge + 1];
        Real a6[6], b6[6];
        for (int k = 0; k < dw; k++) { a6[k] = params[(long long)iw * dw + k]; b6[k] = params[(long long)jw * dw + k]; }
        winterp_begin<Real, M, kMaxJ>(c, a6, b6, W);
        P.qa = q_kept + (long long)iw * n * Nqp + a; P.qb = q_kept + (long long)jw * n * Nqp + b; P.sa = Nqp; P.sb = Nqp;
        flagged++;
        if (!pair_begin<Real, M>(c, eps, P, L, kspin)) verdict = 1;
        else {
          kstart = (k0 >= 2 && k0 <= P.ndivs) ? k0 - 1 : 1;
          P.k = kstart;
          const Real u = P.u_at(P.k);
          lane_begin_lerp<Real, M>(L, P.qa, P.sa, P.qb, P.sb, u);
          winterp_at<Real, M>(W, u, tg);
        }
      }
    }
    if (__all_sync(0xffffffffu, edge < 0)) break;
    const bool running = (edge >= 0 && verdict < 0);
    if (__any_sync(0xffffffffu, running)) {
      PassOut<Real, M> o;
      const bool first = running && P.k == 1;
      if constexpr (PL) lane_pass_planar<Real, true>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      else lane_pass<Real, M, true, SPIN>(c, tg, L, mem, first ? P.qa : mem.prev(), first ? P.sa : mem.T, o);
      const int st = lane_step<Real, M, PL>(c, L, mem, o, running);
      if (running && st != IK_RUNNING) {
        cnt.solves++; cnt.iters += L.it; cnt.evals += L.evals; cnt.ok += (st == IK_OK);
        if (st != IK_OK) verdict = 0;
        else {
          verdict = pair_advance(c, mem, o.s2, thr2, P, L, P.k == kstart && kstart > 1, kspin);
          if (verdict < 0) winterp_at<Real, M>(W, P.u_at(P.k), tg);
        }
      }
    }
    if (edge >= 0 && verdict >= 0) {
      uint32_t *word = &mask[(edge * Nq + a) * Wd + (b >> 5)];
      const uint32_t bit = 1u << (b & 31);
      const uint32_t old = verdict ? atomicOr(word, bit) : atomicAnd(word, ~bit);
      const int was = (old & bit) ? 1 : 0;
      if (was != verdict) {
        flipped++; dvalid += verdict - was;
        if (edge_stats) atomicAdd(&edge_stats[2 * edge + 1], (uint32_t)(verdict - was));
      }
      cnt.pairs++; cnt.valid += verdict;
      edge = -1;
      lane_begin_mem(L, mem);
    }
  }
  if (listed_mode) {
    // explicit pair list (grr_edges_test_listed): these are first tests, counted like the edge kernels count theirs
    cnt.flush(counters);
    if (counters && blockIdx.x == 0 && threadIdx.x == 0 && listed > fcap) atomicAdd(counters + GRR_CNT_FLAG_DROPPED, listed - fcap);
  } else if (counters) {
    warp_add(counters, GRR_CNT_FLAGGED, flagged);
    warp_add(counters, GRR_CNT_FLIPPED, flipped);
    // valid-pair count of the first pass corrected by the verdicts that changed (two's complement add)
    for (int o = 16; o > 0; o >>= 1) dvalid += __shfl_down_sync(0xffffffffu, dvalid, o);
    if ((threadIdx.x & 31) == 0 && dvalid) atomicAdd(counters + GRR_CNT_VALID, (unsigned long long)dvalid);
    if (blockIdx.x == 0 && threadIdx.x == 0 && listed > fcap) atomicAdd(counters + GRR_CNT_FLAG_DROPPED, listed - fcap);
  }
}

// pair-test load of an edge list: out[0] += pairs that will be tested, out[1] += edges with at least one
static __global__ void k_edge_load(long long E, const int32_t *__restrict__ wedges, const int32_t *__restrict__ count,
                            const int32_t *__restrict__ old_count, unsigned long long *out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long p = 0;
  if (e < E) {
    const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
    p = (unsigned long long)count[iw] * (unsigned long long)count[jw];
    if (old_count) p -= (unsigned long long)old_count[iw] * (unsigned long long)old_count[jw];
  }
  unsigned nz = __ballot_sync(0xffffffffu, p != 0);
  for (int o = 16; o > 0; o >>= 1) p += __shfl_down_sync(0xffffffffu, p, o);
  if ((threadIdx.x & 31) == 0 && p) { atomicAdd(out, p); atomicAdd(out + 1, (unsigned long long)__popc(nz)); }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
inline int check_launch(const char *what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -2; }
  return 0;
}

inline int sm_count(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) device = 0;
  if (!cached[device]) cudaDeviceGetAttribute(&cached[device], cudaDevAttrMultiProcessorCount, device);
  return cached[device] > 0 ? cached[device] : 148;
}

constexpr size_t kSmemBudget = 226 * 1024;   // per SM, leaving the static part and the reserved 1 KB

template <typename Kern>
inline int set_smem(Kern kern, size_t smem, const char *what) {
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("%s: dynamic shared memory %zu B: %s", what, smem, cudaGetErrorString(err)); return -2; }
  }
  return 0;
}

inline bool is_planar_y(const HostChain &h) { return is_planar_y_chain(h); }

template <typename Real, int M>
struct Impl {
  using Chain = ChainDev<Real, kMaxJ>;
  static constexpr bool kHasPlanar = (M == 3);
  static bool planar(const HostChain &h) {
    static const bool off = getenv("GRR_NO_PLANAR") != nullptr;   // A/B switch for profiling
    return kHasPlanar && !off && is_planar_y(h);
  }
  static int mc(const HostChain &h) { return planar(h) ? 2 : M; }
  // chains with 'spin' joints run the SPIN instantiations of the kernels (never planar: is_planar_y_chain)
  static bool spin(const HostChain &h) { return spin_mask_of(h) != 0; }

  // persistent lanes: threads per CTA and CTAs so that every SM holds as many lanes as shared memory allows
  static void lane_geometry(const HostChain &h, size_t fixed_bytes, int max_threads, int &threads, int &ctas_per_sm) {
    const size_t per_lane = (size_t)h.n * (4 + mc(h)) * sizeof(Real);
    static const int force_one = getenv("GRR_ONE_CTA_PER_SM") != nullptr;   // experiment switch
    ctas_per_sm = force_one ? 1 : 2;
    // per CTA: 1 KB reserved by the driver + the static shared memory of the kernels (k_edges_build: 4.1 KB of
    // per-warp pair buffers and counters)
    constexpr size_t kPerCtaOverhead = 1024 + 4352;
    size_t per_cta = kSmemBudget / ctas_per_sm - kPerCtaOverhead;
    long t = per_cta > fixed_bytes ? (long)((per_cta - fixed_bytes) / per_lane) : 0;
    if (t < 64) { ctas_per_sm = 1; per_cta = kSmemBudget - kPerCtaOverhead; t = per_cta > fixed_bytes ? (long)((per_cta - fixed_bytes) / per_lane) : 0; }
    if (t > max_threads) t = max_threads;
    threads = (int)(t / 32) * 32;
  }

  static int ik_instances(int mode, const HostChain &h, long long total, int Nq, int Nqp, const void *params,
                          const int32_t *uid, unsigned long long seed, int sample_offset, void *q, uint8_t *status,
                          uint8_t *iters, void *resid, unsigned long long *work, unsigned long long *counters,
                          cudaStream_t st) {
    if (total <= 0) return 0;
    Chain c; fill_chain(h, c);
    int threads, per_sm;
    lane_geometry(h, 0, (sizeof(Real) == 8 && M == 3) ? 192 : 256, threads, per_sm);
    if (threads < 32) { set_error("ik: chain too large for shared memory"); return -2; }
    const size_t smem = (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
    long long ctas = (long long)sm_count(h.device) * per_sm;
    const long long need = (total + threads - 1) / threads;
    if (ctas > need) ctas = need;
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_ik_instances")) return -2;
      kern<<<(unsigned)ctas, threads, smem, st>>>(c, total, Nq, Nqp, (const Real *)params, uid, seed, sample_offset, (Real *)q,
                                                  status, iters, (Real *)resid, work, counters);
      return 0;
    };
    int rc;
    if constexpr (kHasPlanar) {
      if (planar(h)) rc = (mode == 0) ? launch(k_ik_instances<Real, M, 0, true>) : launch(k_ik_instances<Real, M, 1, true>);
      else if (spin(h)) rc = (mode == 0) ? launch(k_ik_instances<Real, M, 0, false, true>) : launch(k_ik_instances<Real, M, 1, false, true>);
      else rc = (mode == 0) ? launch(k_ik_instances<Real, M, 0, false>) : launch(k_ik_instances<Real, M, 1, false>);
    } else if (spin(h)) rc = (mode == 0) ? launch(k_ik_instances<Real, M, 0, false, true>) : launch(k_ik_instances<Real, M, 1, false, true>);
    else rc = (mode == 0) ? launch(k_ik_instances<Real, M, 0, false>) : launch(k_ik_instances<Real, M, 1, false>);
    if (rc) return rc;
    return check_launch("k_ik_instances");
  }

  static int ik_sample(const HostChain &h, long long Nw, int Nq, int Nqp, const void *params, const int32_t *uid,
                       unsigned long long seed, int sample_offset, void *q_raw, uint8_t *status, uint8_t *iters,
                       unsigned long long *work, unsigned long long *counters, cudaStream_t st) {
    return ik_instances(0, h, Nw * (long long)Nq, Nq, Nqp, params, uid, seed, sample_offset, q_raw, status, iters, nullptr,
                        work, counters, st);
  }
  static int ik_level(const HostChain &h, long long Lnodes, const int32_t *node_list, int Nq, int Nqr,
                      const void *params, const int32_t *uid, const int32_t *lower_ptr, const int32_t *lower_idx,
                      const int32_t *degree, const void *Qk, int Nqk, const int32_t *count, unsigned long long seed,
                      int sample_offset, int seed_from_adjacent, void *q_raw, uint8_t *status, int32_t *nadj,
                      unsigned long long *work, unsigned long long *counters, cudaStream_t st) {
    const long long total = Lnodes * (long long)Nq;
    if (total <= 0) return 0;
    Chain c; fill_chain(h, c);
    int threads, per_sm;
    lane_geometry(h, 0, 256, threads, per_sm);
    if (threads < 32) { set_error("ik_level: chain too large for shared memory"); return -2; }
    // a level is small (tens of nodes): spread it over many small CTAs so that every SM gets lanes
    if (threads > 64) threads = 64;
    const size_t smem = (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
    long long ctas = (long long)sm_count(h.device) * 4;
    const long long need = (total + threads - 1) / threads;
    if (ctas > need) ctas = need;
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_ik_level")) return -2;
      kern<<<(unsigned)ctas, threads, smem, st>>>(c, total, node_list, Nq, Nqr, (const Real *)params, uid, lower_ptr,
                                                  lower_idx, degree, (const Real *)Qk, Nqk, count, seed, sample_offset,
                                                  seed_from_adjacent, (Real *)q_raw, status, nadj, work, counters);
      return 0;
    };
    int rc;
    if (spin(h)) rc = launch(k_ik_level<Real, M, false, true>);
    else if constexpr (kHasPlanar) rc = planar(h) ? launch(k_ik_level<Real, M, true>) : launch(k_ik_level<Real, M, false>);
    else rc = launch(k_ik_level<Real, M, false>);
    if (rc) return rc;
    return check_launch("k_ik_level");
  }
  // the whole seeded sweep as one dataflow launch (see k_sweep_dataflow); returns -5 if a node's working set does not
  // fit shared memory (the caller falls back to the level-by-level launches)
  static int ik_sweep(const HostChain &h, long long Nn, const int32_t *order, int Nq, int Nqr, const void *params,
                      const int32_t *uid, const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree, void *Qk,
                      int Nqk, int32_t *count, unsigned long long seed, int sample_offset, int seed_from_adjacent,
                      const void *q_rand, const uint8_t *st_rand, double thresh, int32_t *ready, unsigned long long *work,
                      unsigned long long *counters, cudaStream_t st, int biased, long long start_total, int nodes_with_configs,
                      long long *total_after) {
    if (Nn <= 0) return 0;
    Chain c; fill_chain(h, c);
    const int NqS = (Nq + 3) & ~3, Cmax = (biased && 2 * Nq < 128) ? 128 : 2 * Nq, Wc = (Cmax + 31) / 32;
    if (Wc > 128) return -5;
    int threads = Nq > 128 ? 128 : 64;
    // per lane: the lane machine's joint vectors (+ one result column of the random-start rounds of the biased sweep)
    const size_t per_lane = (size_t)h.n * (4 + mc(h)) * sizeof(Real) + (biased ? (size_t)h.n * sizeof(Real) + 1 : 0);
    const size_t fixed = (size_t)h.n * NqS * sizeof(Real) + (size_t)Cmax * Wc * 4 + (size_t)Cmax * 8 + NqS + Cmax + 64;
    while (threads > 32 && fixed + threads * per_lane > 110 * 1024) threads -= 32;
    const size_t smem = fixed + threads * per_lane;
    if (smem > kSmemBudget - 2048) return -5;
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(ready, 0, sizeof(int32_t) * (size_t)Nn, st);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_sweep_dataflow")) return -2;
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) return -5;
      long long ctas = (long long)sm_count(h.device) * per_sm;          // persistent: every CTA must be resident (it may wait)
      if (ctas > Nn) ctas = Nn;
      kern<<<(unsigned)ctas, threads, smem, st>>>(c, Nn, order, Nq, Nqr, (const Real *)params, uid, lower_ptr, lower_idx, degree,
                                                  (Real *)Qk, Nqk, count, seed, sample_offset, seed_from_adjacent,
                                                  (const Real *)q_rand, st_rand, (Real)thresh, ready, work, counters, start_total,
                                                  nodes_with_configs, total_after);
      return 0;
    };
    int rc;
    if constexpr (kHasPlanar) {
      if (planar(h)) rc = biased ? launch(k_sweep_dataflow<Real, M, true, true>) : launch(k_sweep_dataflow<Real, M, true, false>);
      else if (spin(h)) rc = biased ? launch(k_sweep_dataflow<Real, M, false, true, true>) : launch(k_sweep_dataflow<Real, M, false, false, true>);
      else rc = biased ? launch(k_sweep_dataflow<Real, M, false, true>) : launch(k_sweep_dataflow<Real, M, false, false>);
    } else if (spin(h)) rc = biased ? launch(k_sweep_dataflow<Real, M, false, true, true>) : launch(k_sweep_dataflow<Real, M, false, false, true>);
    else rc = biased ? launch(k_sweep_dataflow<Real, M, false, true>) : launch(k_sweep_dataflow<Real, M, false, false>);
    if (rc) return rc;
    return check_launch("k_sweep_dataflow");
  }
  static int ik_solve_batch(const HostChain &h, long long P, const void *targets, void *q, uint8_t *status,
                            uint8_t *iters, void *resid, unsigned long long *work, unsigned long long *counters,
                            cudaStream_t st) {
    return ik_instances(1, h, P, 1, 1, targets, nullptr, 0, 0, q, status, iters, resid, work, counters, st);
  }
  static int valid_edge_batch(const HostChain &h, long long P, const void *qa, const void *qb, const void *wa,
                              const void *wb, double epsilon, uint8_t *valid, int32_t *info, unsigned long long *work,
                              unsigned long long *counters, cudaStream_t st) {
    if (P <= 0) return 0;
    Chain c; fill_chain(h, c);
    int threads, per_sm;
    lane_geometry(h, 0, 256, threads, per_sm);
    if (threads < 32) { set_error("valid_edge: chain too large for shared memory"); return -2; }
    const size_t smem = (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
    long long ctas = (long long)sm_count(h.device) * per_sm;
    const long long need = (P + threads - 1) / threads;
    if (ctas > need) ctas = need;
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_valid_edge_batch")) return -2;
      kern<<<(unsigned)ctas, threads, smem, st>>>(c, P, (const Real *)qa, (const Real *)qb, (const Real *)wa, (const Real *)wb,
                                                  (Real)epsilon, valid, info, work, counters);
      return 0;
    };
    int rc;
    if (spin(h)) rc = launch(k_valid_edge_batch<Real, M, false, true>);
    else if constexpr (kHasPlanar) rc = planar(h) ? launch(k_valid_edge_batch<Real, M, true>) : launch(k_valid_edge_batch<Real, M, false>);
    else rc = launch(k_valid_edge_batch<Real, M, false>);
    if (rc) return rc;
    return check_launch("k_valid_edge_batch");
  }
  // ex != nullptr: exact build -- single-precision pass with near-tie flags, then the listed pairs again in double
  static int edges_build(const HostChain &h, long long E, const int32_t *wedges, const void *params, const void *q_kept,
                         const int32_t *count, const int32_t *old_count, int Nq, int Nqp, double epsilon,
                         uint32_t *mask, uint32_t *edge_stats, unsigned long long *work, const ScratchReq *sreq,
                         unsigned long long *counters, cudaStream_t st, const ExactArgs *ex) {
    if (E <= 0) return 0;
    if (ex && sizeof(Real) != 4) { set_error("edges_build: the exact build is a single-precision pass + double-precision recheck"); return -1; }
    if (ex && !sreq) { set_error("edges_build: the exact build needs the handle's scratch"); return -1; }
    Chain c; fill_chain(h, c);
    const int Wd = (Nq + 31) / 32;
    const size_t fixed = 2 * (size_t)h.n * Nqp * sizeof(Real) + (size_t)((Nq * Wd + 3) & ~3) * sizeof(uint32_t);
    int threads, per_sm;
    lane_geometry(h, fixed, sizeof(Real) == 4 ? (M == 3 ? 512 : 384) : 256, threads, per_sm);
    // The work per workspace edge decides the kernel: one tiny reduction + a 16-byte readback.
    double avg = 0;
    unsigned long long pairs = 0;
    {
      unsigned long long hload[2] = {0, 0};
      cudaMemsetAsync(work, 0, 2 * sizeof(unsigned long long), st);
      k_edge_load<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(E, wedges, count, old_count, work);
      cudaMemcpyAsync(hload, work, sizeof hload, cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      pairs = hload[0];
      if (hload[1] > 0) avg = (double)hload[0] / (double)hload[1];
    }
    // device scratch: E + 1 queue offsets (global-queue kernel), then the near-tie list (count word + pad + entries)
    unsigned long long *offs = nullptr, *flist = nullptr;
    unsigned long long fcap = 0;
    if (sreq) {
      size_t words = (size_t)E + 2;
      if (ex) { fcap = pairs / 4 + 65536; words += 2 + 2 * (size_t)fcap; }
      offs = sreq->get(sreq->ctx, words);
      if (!offs) return -3;
      if (ex) { flist = offs + ((E + 2) & ~1ll); cudaMemsetAsync(flist, 0, 2 * sizeof(unsigned long long), st); }
    }
    // dense: a CTA owns one workspace edge, so lanes beyond pairs/~6 would idle for the whole edge
    int dense_threads = threads;
    if (avg > 0 && threads >= 32) {
      int want = (int)((avg / 6.0 + 31.0) / 32.0) * 32;
      if (want < 32) want = 32;
      if (want < dense_threads) dense_threads = want;
    }
    // Which kernel: the per-edge CTAs lose the tail of every edge (a lane runs avg/threads pairs whose
    // durations differ by +-50 %), the global queue loses the staging of the node blocks (configs are read from
    // L1/L2, which costs more the longer the chain).  Fitted to the five BASELINE configs (profiles/r01_edge_kernel.md):
    //   dense ~ rounds / (rounds + 1.3),  flat ~ 1 - 0.009 n.
    bool flat = false;
    if (offs) {
      if (threads < 32) flat = true;                              // node blocks too large to stage
      else if (avg > 0) {
        const double rounds = avg / dense_threads;
        flat = (1.0 - 0.009 * h.n) > rounds / (rounds + 1.3);
      }
    }
    const char *force = getenv("GRR_EDGES");                      // "flat" / "dense": A/B switch for tests and profiling
    if (force && offs && !strcmp(force, "flat")) flat = true;
    if (force && threads >= 32 && !strcmp(force, "dense")) flat = false;
    int rc = 0;
    if (flat) rc = edges_flat(h, c, E, wedges, params, q_kept, count, old_count, Nq, Nqp, epsilon, mask, edge_stats, work,
                              offs, counters, st, flist, fcap);
    else {
      if (threads < 32) { set_error("edges_build: Nq=%d, n=%d does not fit shared memory", Nq, h.n); return -2; }
      threads = dense_threads;
      const size_t smem = fixed + (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
      auto launch = [&](auto kern) -> int {
        if (set_smem(kern, smem, "k_edges_build")) return -2;
        kern<<<(unsigned)E, threads, smem, st>>>(c, wedges, (const Real *)params, (const Real *)q_kept, count, old_count, Nq, Nqp,
                                                 (Real)epsilon, mask, edge_stats, counters, flist, fcap);
        return 0;
      };
      const bool pl = planar(h);
      if (spin(h)) {
        if (flist) { set_error("edges_build: the single-precision pass with near-tie flags is not built for chains with spin joints"); return -1; }
        rc = launch(k_edges_build<Real, M, false, false, true>);
      } else
      if constexpr (kHasPlanar && sizeof(Real) == 4) {
        // two lane machines per thread on the packed FP32 instructions (k_edges_build2): planar small-angle chains
        static const bool no_dual = getenv("GRR_DUAL") == nullptr;      // opt-in until validated on the GPU
        static const bool no_kt = getenv("GRR_NO_KT") != nullptr;       // A/B switch: run-time lane stride in the planar pass
        if (pl && c.small_angles != 0 && c.axes_positive != 0 && !no_dual && threads >= 64) {
          const int threads2 = (threads / 2) / 32 * 32;
          const size_t smem2 = fixed + (size_t)threads2 * 2 * h.n * (4 + mc(h)) * sizeof(Real);
          auto launch2 = [&](auto kern) -> int {
            if (set_smem(kern, smem2, "k_edges_build2")) return -2;
            kern<<<(unsigned)E, threads2, smem2, st>>>(c, wedges, (const Real *)params, (const Real *)q_kept, count, old_count, Nq, Nqp,
                                                       (Real)epsilon, mask, edge_stats, counters, flist, fcap);
            return 0;
          };
          rc = flist ? launch2(k_edges_build2<true>) : launch2(k_edges_build2<false>);
        } else
        if (pl && threads == 192 && !no_kt) rc = flist ? launch(k_edges_build<Real, M, true, true, false, 192>) : launch(k_edges_build<Real, M, true, false, false, 192>);
        else if (pl) rc = flist ? launch(k_edges_build<Real, M, true, true>) : launch(k_edges_build<Real, M, true, false>);
        else rc = flist ? launch(k_edges_build<Real, M, false, true>) : launch(k_edges_build<Real, M, false, false>);
      } else if constexpr (kHasPlanar) {
        rc = pl ? launch(k_edges_build<Real, M, true, false>) : launch(k_edges_build<Real, M, false, false>);
      } else if constexpr (sizeof(Real) == 4) {
        rc = flist ? launch(k_edges_build<Real, M, false, true>) : launch(k_edges_build<Real, M, false, false>);
      } else rc = launch(k_edges_build<Real, M, false, false>);
      if (rc) return rc;
      rc = check_launch("k_edges_build");
    }
    if (rc || !ex) return rc;
    if (getenv("GRR_NO_RECHECK")) return rc;                      // timing experiments only: first pass without the second
    return Impl<double, M>::edges_recheck(h, wedges, ex->params64, ex->q64, Nq, Nqp, epsilon, mask, edge_stats, flist, fcap,
                                          work, counters, st);
  }

  // second pass of the exact build: the listed near-tie pairs in this Impl's arithmetic (instantiated for double)
  static int edges_recheck(const HostChain &h, const int32_t *wedges, const void *params, const void *q_kept, int Nq, int Nqp,
                           double epsilon, uint32_t *mask, uint32_t *edge_stats, const unsigned long long *flist,
                           unsigned long long fcap, unsigned long long *work, unsigned long long *counters, cudaStream_t st,
                           int listed_mode = 0) {
    Chain c; fill_chain(h, c);
    // as many lanes per SM as shared memory allows, in one or two CTAs
    int threads, per_sm_want;
    lane_geometry(h, 0, 256, threads, per_sm_want);
    {
      // this kernel has no static shared memory and no staged node blocks: one CTA per SM can hold more lanes than two
      // (planar 20R in double precision: 224 against 2 x 96; the kernel is latency-bound at 6 warps per SM)
      const size_t per_lane = (size_t)h.n * (4 + mc(h)) * sizeof(Real);
      long t1 = (long)((kSmemBudget - 1024) / per_lane) / 32 * 32;
      if (t1 > 256) t1 = 256;
      static const bool off = getenv("GRR_RECHECK_TWO_CTAS") != nullptr;      // A/B switch
      if (!off && t1 > (long)threads * per_sm_want) threads = (int)t1;
    }
    if (threads < 32) { set_error("k_edges_recheck: n=%d does not fit shared memory", h.n); return -2; }
    const size_t smem = (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_edges_recheck")) return -2;
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        set_error("k_edges_recheck: n=%d does not fit shared memory", h.n); return -2;
      }
      kern<<<(unsigned)(sm_count(h.device) * per_sm), threads, smem, st>>>(c, wedges, (const Real *)params, (const Real *)q_kept, Nq,
                                                                           Nqp, (Real)epsilon, mask, edge_stats, flist, fcap, work,
                                                                           counters, listed_mode);
      return 0;
    };
    int rc;
    if (spin(h)) rc = launch(k_edges_recheck<Real, M, false, true>);
    else if constexpr (kHasPlanar) rc = planar(h) ? launch(k_edges_recheck<Real, M, true>) : launch(k_edges_recheck<Real, M, false>);
    else rc = launch(k_edges_recheck<Real, M, false>);
    if (rc) return rc;
    return check_launch("k_edges_recheck");
  }

  // explicit pair list {edge << 32 | a << 16 | b, 0}: each listed pair is tested and its mask bit set or cleared
  // (the tests of the visibility-graph variant: knn pairs of constructSelfMotionManifold, testAndConnectQccs candidates)
  static int edges_listed(const HostChain &h, const int32_t *wedges, const void *params, const void *q_kept, int Nq, int Nqp,
                          double epsilon, uint32_t *mask, uint32_t *edge_stats, const unsigned long long *list,
                          unsigned long long cap, unsigned long long *work, unsigned long long *counters, cudaStream_t st) {
    return edges_recheck(h, wedges, params, q_kept, Nq, Nqp, epsilon, mask, edge_stats, list, cap, work, counters, st, 1);
  }

  static int edges_flat(const HostChain &h, const Chain &c, long long E, const int32_t *wedges, const void *params,
                        const void *q_kept, const int32_t *count, const int32_t *old_count, int Nq, int Nqp, double epsilon,
                        uint32_t *mask, uint32_t *edge_stats, unsigned long long *work, unsigned long long *offs,
                        unsigned long long *counters, cudaStream_t st, unsigned long long *flist, unsigned long long fcap) {
    const int Wd = (Nq + 31) / 32;
    k_edges_prepare<<<(unsigned)E, 128, 0, st>>>(E, wedges, count, old_count, Nq, Wd, mask, offs, edge_stats);
    if (int rc = check_launch("k_edges_prepare")) return rc;
    k_scan_u64<<<1, 1024, 0, st>>>(E, offs);
    if (int rc = check_launch("k_scan_u64")) return rc;
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    const int threads = 128;
    static_assert(GRR_FLAT_KT == 0 || GRR_FLAT_KT == 128, "k_edges_flat is launched with 128 lanes per CTA");
    const size_t smem = (size_t)(threads / 32) * 32 * sizeof(int4) + (size_t)threads * h.n * (4 + mc(h)) * sizeof(Real);
    auto launch = [&](auto kern) -> int {
      if (set_smem(kern, smem, "k_edges_flat")) return -2;
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        set_error("k_edges_flat: n=%d does not fit shared memory", h.n); return -2;
      }
      kern<<<(unsigned)(sm_count(h.device) * per_sm), threads, smem, st>>>(c, E, wedges, (const Real *)params, (const Real *)q_kept,
                                                                           count, old_count, Nq, Nqp, (Real)epsilon, mask, offs,
                                                                           work, counters, flist, fcap);
      return 0;
    };
    int rc;
    const bool pl = planar(h);
    if (spin(h)) {
      if (flist) { set_error("edges_build: the single-precision pass with near-tie flags is not built for chains with spin joints"); return -1; }
      rc = launch(k_edges_flat<Real, M, false, false, true>);
    } else
    if constexpr (kHasPlanar && sizeof(Real) == 4) {
      if (pl) rc = flist ? launch(k_edges_flat<Real, M, true, true>) : launch(k_edges_flat<Real, M, true, false>);
      else rc = flist ? launch(k_edges_flat<Real, M, false, true>) : launch(k_edges_flat<Real, M, false, false>);
    } else if constexpr (kHasPlanar) {
      rc = pl ? launch(k_edges_flat<Real, M, true, false>) : launch(k_edges_flat<Real, M, false, false>);
    } else if constexpr (sizeof(Real) == 4) {
      rc = flist ? launch(k_edges_flat<Real, M, false, true>) : launch(k_edges_flat<Real, M, false, false>);
    } else rc = launch(k_edges_flat<Real, M, false, false>);
    if (rc) return rc;
    if (int rc2 = check_launch("k_edges_flat")) return rc2;
    if (edge_stats) {
      k_edges_bits<<<(unsigned)E, 128, 0, st>>>(E, Nq * Wd, mask, edge_stats);
      return check_launch("k_edges_bits");
    }
    return 0;
  }
};

}  // namespace grr
