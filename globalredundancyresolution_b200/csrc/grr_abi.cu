// grr_abi.cu -- the extern "C" boundary of libgrr_b200.so (see include/grr_b200.h) plus the
// kernels that do not depend on the joint count: duplicate filter, bitmask -> edge-list
// compaction, FK, node-block gather/scatter for the halo exchange, FP32 peak microbenchmark.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <dlfcn.h>
#include <nccl.h>
#include <new>
#include "grr_device.cuh"
#include "grr_host.h"
#include "grr_fill.cuh"

namespace grr {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

static inline int launch_ok(const char *what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -2; }
  return 0;
}

// ------------------------------------------------------------------------------------------
// S3: greedy first-wins duplicate filter, one CTA per workspace node
// (solver.py:802-813,824-835,886-897: keep q iff robot.distance(q,q') >= 1e-2 for all kept q')
// ------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128)
k_dedup(int n, unsigned spin, int Nq, int Nqp, const Real *__restrict__ q_raw, const uint8_t *__restrict__ status, Real thresh,
        int Nqp_kept, Real *q_kept, const int32_t *count_in, int32_t *__restrict__ keep_idx,
        int32_t *count, const int32_t *__restrict__ node_list) {
  // raw candidates are indexed by block (level-local when node_list is given), kept configs by workspace node
  const long long blk = blockIdx.x;
  const long long node = node_list ? node_list[blk] : blk;
  const Real *raw = q_raw + blk * n * (long long)Nqp;
  Real *out = q_kept + node * n * (long long)Nqp_kept;
  const uint8_t *st = status + blk * Nqp;
  int kept = count_in ? count_in[node] : 0;
  const int kept0 = kept;
  for (int s = 0; s < Nq; s++) {
    if (st[s] != 0) continue;                     // failed IK: q is None (solver.py:800,822,885)
    int same = 0;
    for (int k = threadIdx.x; k < kept; k += blockDim.x) {
      Real d2 = 0;
      for (int j = 0; j < n; j++) { Real d = joint_diff(spin, j, raw[j * Nqp + s], out[j * Nqp_kept + k]); d2 += d * d; }
      if (sqrt_(d2) < thresh) same = 1;
    }
    same = __syncthreads_or(same);
    if (!same && kept < Nqp_kept) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) out[j * Nqp_kept + kept] = raw[j * Nqp + s];
      if (threadIdx.x == 0 && keep_idx) keep_idx[blk * Nq + (kept - kept0)] = s;
      kept++;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) count[node] = kept;
  if (keep_idx)
    for (int k = (kept - kept0) + threadIdx.x; k < Nq; k += blockDim.x) keep_idx[blk * Nq + k] = -1;
}

// ------------------------------------------------------------------------------------------
// Dedup of one wavefront level of the seeded sweep (solver.py:794-835): candidates in the sweep's order --
// the node's nadj seeded attempts (level-local), then the first Nq - nadj + numFailed precomputed random-start
// attempts of the node -- against the node's kept list, in place.  Counts the attempts the sweep ran.
// ------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128)
k_dedup_sweep(int n, unsigned spin, int Nq, const int32_t *__restrict__ node_list, int Nqs, const Real *__restrict__ q_seed,
              const uint8_t *__restrict__ st_seed, const int32_t *__restrict__ nadj, int use_seeded, int Nqr,
              const Real *__restrict__ q_rand, const uint8_t *__restrict__ st_rand, Real thresh, int Nqp_kept, Real *q_kept,
              int32_t *count, unsigned long long *counters) {
  const long long blk = blockIdx.x;
  const long long node = node_list[blk];
  const Real *seedq = q_seed + blk * n * (long long)Nqs;
  const uint8_t *sts = st_seed + blk * Nqs;
  const Real *randq = q_rand + node * n * (long long)Nqr;
  const uint8_t *str = st_rand + node * Nqr;
  Real *out = q_kept + node * n * (long long)Nqp_kept;
  const int na = nadj[blk];
  int nfail = 0;
  if (use_seeded) for (int s = 0; s < na; s++) nfail += (sts[s] != 0);
  const int nseed = use_seeded ? na : 0, nrand = Nq - na + nfail;
  int kept = count[node], nok = 0;
  for (int cidx = 0; cidx < nseed + nrand; cidx++) {
    const bool sd = cidx < nseed;
    const int s = sd ? cidx : cidx - nseed;
    const uint8_t stv = sd ? sts[s] : str[s];
    if (stv != 0) continue;
    nok++;
    const Real *cand = sd ? seedq + s : randq + s;
    const int cs = sd ? Nqs : Nqr;
    int same = 0;
    for (int k = threadIdx.x; k < kept; k += blockDim.x) {
      Real d2 = 0;
      for (int j = 0; j < n; j++) { Real d = joint_diff(spin, j, cand[(long long)j * cs], out[j * Nqp_kept + k]); d2 += d * d; }
      if (sqrt_(d2) < thresh) same = 1;
    }
    same = __syncthreads_or(same);
    if (!same && kept < Nqp_kept) {
      for (int j = threadIdx.x; j < n; j += blockDim.x) out[j * Nqp_kept + kept] = cand[(long long)j * cs];
      kept++;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    count[node] = kept;
    if (counters) {
      atomicAdd(counters + GRR_CNT_IK_SOLVES, (unsigned long long)(nseed + nrand));
      atomicAdd(counters + GRR_CNT_IK_OK, (unsigned long long)nok);
    }
  }
}

// ------------------------------------------------------------------------------------------
// bitmask -> (iq,jq) lists, one CTA per workspace edge, ballot-free ordered compaction:
// block-wide exclusive scan of the mask words' popcounts, then each thread expands its word.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_edges_compact(int Nq, const uint32_t *__restrict__ mask, const long long *__restrict__ offsets, int32_t *__restrict__ out) {
  const long long e = blockIdx.x;
  const int Wd = (Nq + 31) >> 5, words = Nq * Wd;
  const uint32_t *m = mask + e * (long long)words;
  __shared__ int s_warp[4];
  __shared__ long long s_base;
  if (threadIdx.x == 0) s_base = offsets[e];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w0 = 0; w0 < words; w0 += blockDim.x) {
    const int wi = w0 + threadIdx.x;
    const uint32_t word = wi < words ? m[wi] : 0u;
    const int c = __popc(word);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int k = 0; k < 4; k++) { if (k < warp) wbase += s_warp[k]; total += s_warp[k]; }
    long long pos = s_base + wbase + incl - c;
    if (c) {
      const int a = wi / Wd, b0 = (wi - a * Wd) * 32;
      uint32_t bits = word;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        out[2 * pos] = a; out[2 * pos + 1] = b0 + b;
        pos++;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Work estimate per workspace edge: sum over its (not both-old) pairs of Ndivs + 1 = ceil(dist/eps) + 1
// (graph.py:944-946), i.e. the number of interpolated IK solves + leaf tests the edge kernel will run.
// Used to balance shards across GPUs by work instead of by pair count.  One CTA per workspace edge.
// ------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(256)
k_edges_cost(int n, unsigned spin, int nlinks_eps, const int32_t *__restrict__ wedges, const Real *__restrict__ q_kept,
             const int32_t *__restrict__ count, const int32_t *__restrict__ old_count, int Nqp, Real epsilon,
             unsigned long long *__restrict__ cost) {
  extern __shared__ __align__(16) unsigned char smem_cost[];
  Real *sa = reinterpret_cast<Real *>(smem_cost), *sb = sa + n * Nqp;
  __shared__ unsigned long long s_sum;
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e], jw = wedges[2 * e + 1];
  const int ca = count[iw], cb = count[jw];
  const int oa = old_count ? old_count[iw] : 0, ob = old_count ? old_count[jw] : 0;
  if (threadIdx.x == 0) s_sum = 0;
  const unsigned total = (unsigned)ca * (unsigned)cb;
  if (total) {
    const Real *ga = q_kept + (long long)iw * n * Nqp, *gb = q_kept + (long long)jw * n * Nqp;
    for (int k = threadIdx.x; k < n * Nqp; k += blockDim.x) { sa[k] = ga[k]; sb[k] = gb[k]; }
  }
  __syncthreads();
  const Real eps = epsilon * sqrt_(Real(nlinks_eps));
  unsigned long long acc = 0;
  for (unsigned t = threadIdx.x; t < total; t += blockDim.x) {
    const int a = (int)(t / (unsigned)cb), b = (int)(t - (unsigned)a * (unsigned)cb);
    if (a < oa && b < ob) continue;
    Real d2 = 0;
    for (int j = 0; j < n; j++) { const Real d = joint_diff(spin, j, sa[j * Nqp + a], sb[j * Nqp + b]); d2 += d * d; }
    acc += (unsigned long long)ceil_(sqrt_(d2) / eps) + 1ull;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&s_sum, acc);
  __syncthreads();
  if (threadIdx.x == 0) cost[e] = s_sum;
}

// ------------------------------------------------------------------------------------------
// FK on a batch (API: ikError / getIKParameters; graph.py:747-756,887-893)
// ------------------------------------------------------------------------------------------
template <typename Real>
__global__ void k_fk_batch(const __grid_constant__ ChainDev<Real, kMaxJ> c, long long P, const Real *__restrict__ q,
                           Real *__restrict__ out_p, Real *__restrict__ out_R) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P) return;
  Real x[kMaxJ], p[3], R[9];
  for (int j = 0; j < c.n; j++) x[j] = q[(long long)j * P + t];
  chain_fk<Real, kMaxJ>(c, x, p, R);
  out_p[3 * t] = p[0]; out_p[3 * t + 1] = p[1]; out_p[3 * t + 2] = p[2];
  if (out_R) for (int k = 0; k < 9; k++) out_R[9 * t + k] = R[k];
}

// ------------------------------------------------------------------------------------------
// halo exchange support: whole node blocks <-> packed message buffer
// ------------------------------------------------------------------------------------------
template <typename Real, bool GATHER>
__global__ void k_node_blocks(long long K, const int32_t *__restrict__ idx, int blk, Real *q, int32_t *count, Real *packed,
                              int32_t *packed_count) {
  const long long k = blockIdx.x;
  const long long node = idx[k];
  Real *a = q + node * blk, *b = packed + k * blk;
  for (int i = threadIdx.x; i < blk; i += blockDim.x) { if (GATHER) b[i] = a[i]; else a[i] = b[i]; }
  if (threadIdx.x == 0) { if (GATHER) packed_count[k] = count[node]; else count[node] = packed_count[k]; }
}

// ------------------------------------------------------------------------------------------
// FP32 FMA peak (roofline denominator for the IK / edge kernels)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// double precision FMA peak (roofline denominator of the double-precision kernels: sampling, general-chain edge phase)
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// packed variant: sm_100 FFMA2 (two FP32 FMAs per instruction per lane)
__global__ void __launch_bounds__(256) k_fma2_peak(float *out, int iters, float a, float b) {
  float2 x0 = make_float2(threadIdx.x, threadIdx.x + 1.f), x1 = make_float2(x0.x + 2.f, x0.x + 3.f);
  float2 x2 = make_float2(x0.x + 4.f, x0.x + 5.f), x3 = make_float2(x0.x + 6.f, x0.x + 7.f);
  float2 x4 = make_float2(x0.x + 8.f, x0.x + 9.f), x5 = make_float2(x0.x + 10.f, x0.x + 11.f);
  float2 x6 = make_float2(x0.x + 12.f, x0.x + 13.f), x7 = make_float2(x0.x + 14.f, x0.x + 15.f);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B); x2 = __ffma2_rn(x2, A, B); x3 = __ffma2_rn(x3, A, B);
      x4 = __ffma2_rn(x4, A, B); x5 = __ffma2_rn(x5, A, B); x6 = __ffma2_rn(x6, A, B); x7 = __ffma2_rn(x7, A, B);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + x4.x + x4.y + x5.x + x5.y +
                                               x6.x + x6.y + x7.x + x7.y;
}

static const Ops *pick_ops(int dtype, int mode) {
  const bool m6 = (mode != GRR_FREE);
  if (dtype == GRR_F32) return m6 ? ops_f32_m6() : ops_f32_m3();
  if (dtype == GRR_F64) return m6 ? ops_f64_m6() : ops_f64_m3();
  return nullptr;
}

}  // namespace grr

using namespace grr;

struct grr_handle_s {
  HostChain hc;
  unsigned long long *d_work;   // ring of device work counters for the persistent-lane kernels
  unsigned next_work;
  unsigned long long *d_scratch;   // grow-only device scratch (pair-queue offsets of grr_edges_build)
  size_t scratch_words;
};
static const unsigned kWorkRing = 64;

// next work counter of the handle's ring (allocated on first use, on the handle's device)
static unsigned long long *work_slot(grr_handle_s *h) {
  if (!h->d_work) {
    if (cudaMalloc(&h->d_work, sizeof(unsigned long long) * 2 * kWorkRing) != cudaSuccess) {
      set_error("cudaMalloc(work counters) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
  }
  // two words per slot (grr_edges_build's load reduction uses both); one handle serves one stream at a time
  return h->d_work + 2 * (h->next_work++ % kWorkRing);
}

// device scratch of at least `words` 64-bit words (grow-only; reallocation synchronises the device)
static unsigned long long *scratch_for(grr_handle_s *h, size_t words) {
  if (h->scratch_words < words) {
    if (h->d_scratch) { cudaDeviceSynchronize(); cudaFree(h->d_scratch); h->d_scratch = nullptr; h->scratch_words = 0; }
    const size_t want = words + words / 4 + 1024;
    if (cudaMalloc(&h->d_scratch, sizeof(unsigned long long) * want) != cudaSuccess) {
      set_error("cudaMalloc(scratch, %zu words) failed: %s", want, cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    h->scratch_words = want;
  }
  return h->d_scratch;
}

static unsigned long long *scratch_cb(void *ctx, size_t words) { return scratch_for((grr_handle_s *)ctx, words); }

#define GRR_CHECK_HANDLE(h)                                   \
  if (!(h)) { set_error("null handle"); return -1; }          \
  { cudaError_t e_ = cudaSetDevice((h)->hc.device);           \
    if (e_ != cudaSuccess) { set_error("cudaSetDevice(%d): %s", (h)->hc.device, cudaGetErrorString(e_)); return -2; } }

extern "C" {

const char *grr_last_error(void) { return g_err; }
int grr_abi_version(void) { return 2; }   // 2: spin joints (grr_chain_set_spin_joints, spin_mask of the visibility entry points)

int grr_chain_create(const grr_chain_desc *d, int device, grr_handle_t *out) {
  if (!d || !out) { set_error("grr_chain_create: null argument"); return -1; }
  if (d->n < 1 || d->n > GRR_MAX_JOINTS) { set_error("grr_chain_create: n=%d outside 1..%d", d->n, GRR_MAX_JOINTS); return -1; }
  if (d->mode < GRR_FREE || d->mode > GRR_VARIABLE) { set_error("grr_chain_create: bad mode %d", d->mode); return -1; }
  if (!d->Rp || !d->tp || !d->axis || !d->qmin || !d->qmax) { set_error("grr_chain_create: null chain arrays"); return -1; }
  if (!(d->tol > 0) || d->max_iters < 0 || d->max_iters > 250 || !(d->lambda >= 0)) { set_error("grr_chain_create: bad solver parameters"); return -1; }
  if (d->nlinks_eps < 1) { set_error("grr_chain_create: nlinks_eps must be >= 1"); return -1; }
  grr_handle_s *h = new (std::nothrow) grr_handle_s;
  if (!h) { set_error("out of memory"); return -3; }
  memset(h, 0, sizeof *h);
  HostChain &c = h->hc;
  c.n = d->n; c.mode = d->mode; c.nlinks_eps = d->nlinks_eps; c.max_iters = d->max_iters; c.device = device;
  c.tol = d->tol; c.lambda = d->lambda;
  memcpy(c.Rp, d->Rp, sizeof(double) * 9 * d->n);
  memcpy(c.tp, d->tp, sizeof(double) * 3 * d->n);
  memcpy(c.axis, d->axis, sizeof(double) * 3 * d->n);
  memcpy(c.qmin, d->qmin, sizeof(double) * d->n);
  memcpy(c.qmax, d->qmax, sizeof(double) * d->n);
  memcpy(c.ee_local, d->ee_local, sizeof c.ee_local);
  memcpy(c.R_tail, d->R_tail, sizeof c.R_tail);
  memcpy(c.R_goal, d->R_goal, sizeof c.R_goal);
  for (int j = 0; j < c.n; j++) {
    const double *a = c.axis + 3 * j;
    const double nn = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    if (!(nn > 0.999 && nn < 1.001)) { set_error("grr_chain_create: axis %d is not unit", j); delete h; return -1; }
    if (!(c.qmin[j] <= c.qmax[j])) { set_error("grr_chain_create: qmin>qmax at joint %d", j); delete h; return -1; }
  }
  {
    // near-tie margins of grr_edges_build_exact: safety factor 8 over the measured single-precision differences
    // (DESIGN.md 3.4); GRR_FLAG_SCALE overrides it for experiments
    const char *fs = getenv("GRR_FLAG_SCALE");
    flag_margins(c, fs ? atof(fs) : 8.0);
  }
  *out = h;
  return 0;
}

int grr_chain_set_spin_joints(grr_handle_t h, const int32_t *flags) {
  if (!h) { set_error("null handle"); return -1; }
  for (int j = 0; j < h->hc.n; j++) h->hc.spin[j] = (flags && flags[j]) ? 1 : 0;
  return 0;
}

int grr_set_flag_scale(grr_handle_t h, double scale) {
  if (!h) { set_error("null handle"); return -1; }
  if (!(scale > 0)) { set_error("grr_set_flag_scale: scale must be positive"); return -1; }
  flag_margins(h->hc, scale);
  return 0;
}

int grr_chain_destroy(grr_handle_t h) {
  if (h && (h->d_work || h->d_scratch)) {
    cudaSetDevice(h->hc.device);
    if (h->d_work) cudaFree(h->d_work);
    if (h->d_scratch) cudaFree(h->d_scratch);
  }
  delete h;
  return 0;
}

int grr_task_set(grr_handle_t h, int mode, const double *R_goal9) {
  if (!h) { set_error("null handle"); return -1; }
  if (mode < GRR_FREE || mode > GRR_VARIABLE) { set_error("grr_task_set: bad mode %d", mode); return -1; }
  h->hc.mode = mode;
  if (R_goal9) memcpy(h->hc.R_goal, R_goal9, sizeof h->hc.R_goal);
  return 0;
}

int grr_chain_num_joints(grr_handle_t h) { return h ? h->hc.n : -1; }

int grr_fk_batch(grr_handle_t h, int dtype, int64_t P, const void *q, void *out_p, void *out_R, void *stream) {
  GRR_CHECK_HANDLE(h);
  if (P <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((P + 127) / 128);
  if (dtype == GRR_F32) {
    ChainDev<float, kMaxJ> c;
    fill_chain(h->hc, c);
    k_fk_batch<float><<<blocks, 128, 0, st>>>(c, P, (const float *)q, (float *)out_p, (float *)out_R);
  } else if (dtype == GRR_F64) {
    ChainDev<double, kMaxJ> c;
    fill_chain(h->hc, c);
    k_fk_batch<double><<<blocks, 128, 0, st>>>(c, P, (const double *)q, (double *)out_p, (double *)out_R);
  } else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_fk_batch");
}

int grr_ik_solve_batch(grr_handle_t h, int dtype, int64_t P, const void *targets, void *q, uint8_t *status,
                       uint8_t *iters, void *resid, uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (P < 0 || (P > 0 && (!targets || !q || !status))) { set_error("grr_ik_solve_batch: null buffer"); return -1; }
  if (P == 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  return ops->ik_solve_batch(h->hc, P, targets, q, status, iters, resid, work, (unsigned long long *)counters, (cudaStream_t)stream);
}

int grr_ik_sample(grr_handle_t h, int dtype, int64_t Nw, int32_t Nq, int32_t Nqp, const void *params,
                  const int32_t *node_uid, uint64_t seed, int32_t sample_offset, void *q_raw, uint8_t *status,
                  uint8_t *iters, uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (Nw < 0 || Nq < 0 || Nqp < Nq) { set_error("grr_ik_sample: bad sizes Nw=%lld Nq=%d Nqp=%d", (long long)Nw, Nq, Nqp); return -1; }
  if (Nw * (int64_t)Nq > 0 && (!params || !node_uid || !q_raw || !status)) { set_error("grr_ik_sample: null buffer"); return -1; }
  if (Nw * (int64_t)Nq == 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  return ops->ik_sample(h->hc, Nw, Nq, Nqp, params, node_uid, seed, sample_offset, q_raw, status, iters, work,
                        (unsigned long long *)counters, (cudaStream_t)stream);
}

int grr_ik_sample_level(grr_handle_t h, int dtype, int64_t L, const int32_t *node_list, int32_t Nq, int32_t Nqr,
                        const void *params, const int32_t *node_uid, const int32_t *lower_ptr, const int32_t *lower_idx,
                        const int32_t *degree, const void *q_kept, int32_t Nqp_kept, const int32_t *count, uint64_t seed,
                        int32_t sample_offset, int32_t seed_from_adjacent, void *q_raw, uint8_t *status, int32_t *nadj,
                        uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (L <= 0 || Nq <= 0) return 0;
  if (Nqr < Nq) { set_error("grr_ik_sample_level: Nqr < Nq"); return -1; }
  if (!node_list || !params || !node_uid || !lower_ptr || !lower_idx || !degree || !q_kept || !count || !q_raw || !status || !nadj) {
    set_error("grr_ik_sample_level: null buffer"); return -1;
  }
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  return ops->ik_level(h->hc, L, node_list, Nq, Nqr, params, node_uid, lower_ptr, lower_idx, degree, q_kept, Nqp_kept, count,
                       seed, sample_offset, seed_from_adjacent, q_raw, status, nadj, work, (unsigned long long *)counters,
                       (cudaStream_t)stream);
}

int grr_ik_sample_sweep(grr_handle_t h, int dtype, int64_t Nw, const int32_t *order, int32_t Nq, int32_t Nqr, const void *params,
                        const int32_t *node_uid, const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree,
                        void *q_kept, int32_t Nqp_kept, int32_t *count, uint64_t seed, int32_t sample_offset,
                        int32_t seed_from_adjacent, const void *q_rand, const uint8_t *status_rand, double thresh, int32_t *ready,
                        uint64_t *counters, void *stream, int32_t biased, int64_t start_total, int32_t nodes_with_configs,
                        int64_t *total_after) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (Nw <= 0 || Nq <= 0) return 0;
  if (Nqr < Nq) { set_error("grr_ik_sample_sweep: Nqr < Nq"); return -1; }
  if (!order || !params || !node_uid || !lower_ptr || !lower_idx || !degree || !q_kept || !count || !ready ||
      (!biased && (!q_rand || !status_rand)) || (biased && !total_after)) {
    set_error("grr_ik_sample_sweep: null buffer"); return -1;
  }
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  const int rc = ops->ik_sweep(h->hc, Nw, order, Nq, Nqr, params, node_uid, lower_ptr, lower_idx, degree, q_kept, Nqp_kept, count, seed,
                               sample_offset, seed_from_adjacent, q_rand, status_rand, thresh, ready, work,
                               (unsigned long long *)counters, (cudaStream_t)stream, biased ? 1 : 0, (long long)start_total,
                               nodes_with_configs, (long long *)total_after);
  if (rc == -5) set_error("grr_ik_sample_sweep: Nq=%d, n=%d: a node's working set does not fit shared memory (use the level launches)", Nq, h->hc.n);
  return rc;
}

int grr_dedup(grr_handle_t h, int dtype, int64_t Nw, int32_t Nq, int32_t Nqp, const void *q_raw, const uint8_t *status,
              double thresh, int32_t Nqp_kept, void *q_kept, const int32_t *count_in, int32_t *keep_idx, int32_t *count,
              void *stream) {
  GRR_CHECK_HANDLE(h);
  if (Nw <= 0) return 0;
  if (!q_raw || !status || !q_kept || !count) { set_error("grr_dedup: null buffer"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GRR_F32)
    k_dedup<float><<<(unsigned)Nw, 128, 0, st>>>(h->hc.n, spin_mask_of(h->hc), Nq, Nqp, (const float *)q_raw, status, (float)thresh, Nqp_kept,
                                                 (float *)q_kept, count_in, keep_idx, count, nullptr);
  else if (dtype == GRR_F64)
    k_dedup<double><<<(unsigned)Nw, 128, 0, st>>>(h->hc.n, spin_mask_of(h->hc), Nq, Nqp, (const double *)q_raw, status, thresh, Nqp_kept,
                                                  (double *)q_kept, count_in, keep_idx, count, nullptr);
  else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_dedup");
}

int grr_dedup_sweep(grr_handle_t h, int dtype, int64_t L, const int32_t *node_list, int32_t Nq, int32_t Nqs, const void *q_seed,
                    const uint8_t *status_seed, const int32_t *nadj, int32_t seed_from_adjacent, int32_t Nqr, const void *q_rand,
                    const uint8_t *status_rand, double thresh, int32_t Nqp_kept, void *q_kept, int32_t *count,
                    uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  if (L <= 0) return 0;
  if (!node_list || !q_seed || !status_seed || !nadj || !q_rand || !status_rand || !q_kept || !count) {
    set_error("grr_dedup_sweep: null buffer"); return -1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GRR_F32)
    k_dedup_sweep<float><<<(unsigned)L, 128, 0, st>>>(h->hc.n, spin_mask_of(h->hc), Nq, node_list, Nqs, (const float *)q_seed, status_seed, nadj,
                                                      seed_from_adjacent, Nqr, (const float *)q_rand, status_rand, (float)thresh,
                                                      Nqp_kept, (float *)q_kept, count, (unsigned long long *)counters);
  else if (dtype == GRR_F64)
    k_dedup_sweep<double><<<(unsigned)L, 128, 0, st>>>(h->hc.n, spin_mask_of(h->hc), Nq, node_list, Nqs, (const double *)q_seed, status_seed, nadj,
                                                       seed_from_adjacent, Nqr, (const double *)q_rand, status_rand, thresh,
                                                       Nqp_kept, (double *)q_kept, count, (unsigned long long *)counters);
  else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_dedup_sweep");
}

int grr_edges_build(grr_handle_t h, int dtype, int64_t E, const int32_t *wedges, const void *params, const void *q_kept,
                    const int32_t *count, const int32_t *old_count, int32_t Nq, int32_t Nqp, double epsilon,
                    uint32_t *mask, uint32_t *edge_stats, uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (E < 0 || Nq < 1 || Nq > 65535 || Nqp < Nq) { set_error("grr_edges_build: bad sizes (need 1 <= Nq <= 65535, Nqp >= Nq)"); return -1; }
  if (E > 0 && (!wedges || !params || !q_kept || !count || !mask)) { set_error("grr_edges_build: null buffer"); return -1; }
  if (E == 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  const ScratchReq sreq = {h, scratch_cb};
  return ops->edges_build(h->hc, E, wedges, params, q_kept, count, old_count, Nq, Nqp, epsilon, mask, edge_stats, work, &sreq,
                          (unsigned long long *)counters, (cudaStream_t)stream, nullptr);
}

int grr_edges_test_listed(grr_handle_t h, int dtype, const int32_t *wedges, const void *params, const void *q_kept, int32_t Nq,
                          int32_t Nqp, double epsilon, uint32_t *mask, uint32_t *edge_stats, const uint64_t *list, uint64_t cap,
                          uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (!wedges || !params || !q_kept || !mask || !list || Nq < 1 || Nqp < Nq || Nq > 65535) { set_error("grr_edges_test_listed: bad argument"); return -1; }
  if (cap == 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  return ops->edges_listed(h->hc, wedges, params, q_kept, Nq, Nqp, epsilon, mask, edge_stats, (const unsigned long long *)list, cap, work,
                           (unsigned long long *)counters, (cudaStream_t)stream);
}

int grr_edges_exact_mode(grr_handle_t h) {
  if (!h) { set_error("null handle"); return -1; }
  const char *force = getenv("GRR_EXACT_MODE");                   // "recheck" / "double": experiments
  if (force && !strcmp(force, "recheck")) return 1;
  if (force && !strcmp(force, "double")) return 2;
  return (h->hc.mode == GRR_FREE && is_planar_y_chain(h->hc)) ? 1 : 2;
}

int grr_edges_build_exact(grr_handle_t h, int64_t E, const int32_t *wedges, const float *params32, const double *params64,
                          const float *q_kept32, const double *q_kept64, const int32_t *count, const int32_t *old_count,
                          int32_t Nq, int32_t Nqp, double epsilon, uint32_t *mask, uint32_t *edge_stats, uint64_t *counters,
                          void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(GRR_F32, h->hc.mode);
  if (E < 0 || Nq < 1 || Nq > 65535 || Nqp < Nq) { set_error("grr_edges_build_exact: bad sizes (need 1 <= Nq <= 65535, Nqp >= Nq)"); return -1; }
  if (E > 0 && (!wedges || !params32 || !params64 || !q_kept32 || !q_kept64 || !count || !mask)) { set_error("grr_edges_build_exact: null buffer"); return -1; }
  if (E == 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  const ScratchReq sreq = {h, scratch_cb};
  if (grr_edges_exact_mode(h) == 2) {
    // general chains: the double-precision kernels on the double-precision configs (see grr_edges_exact_mode)
    const Ops *ops64 = pick_ops(GRR_F64, h->hc.mode);
    return ops64->edges_build(h->hc, E, wedges, params64, q_kept64, count, old_count, Nq, Nqp, epsilon, mask, edge_stats, work, &sreq,
                              (unsigned long long *)counters, (cudaStream_t)stream, nullptr);
  }
  const ExactArgs ex = {params64, q_kept64};
  return ops->edges_build(h->hc, E, wedges, params32, q_kept32, count, old_count, Nq, Nqp, epsilon, mask, edge_stats, work, &sreq,
                          (unsigned long long *)counters, (cudaStream_t)stream, &ex);
}

int grr_edges_cost(grr_handle_t h, int dtype, int64_t E, const int32_t *wedges, const void *q_kept, const int32_t *count,
                   const int32_t *old_count, int32_t Nqp, double epsilon, uint64_t *cost, void *stream) {
  GRR_CHECK_HANDLE(h);
  if (E <= 0) return 0;
  if (!wedges || !q_kept || !count || !cost) { set_error("grr_edges_cost: null buffer"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = h->hc.n;
  const size_t smem = 2 * (size_t)n * Nqp * (dtype == GRR_F64 ? 8 : 4);
  if (dtype == GRR_F32) {
    auto kern = k_edges_cost<float>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)E, 256, smem, st>>>(n, spin_mask_of(h->hc), h->hc.nlinks_eps, wedges, (const float *)q_kept, count, old_count, Nqp, (float)epsilon,
                                         (unsigned long long *)cost);
  } else if (dtype == GRR_F64) {
    auto kern = k_edges_cost<double>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)E, 256, smem, st>>>(n, spin_mask_of(h->hc), h->hc.nlinks_eps, wedges, (const double *)q_kept, count, old_count, Nqp, epsilon,
                                         (unsigned long long *)cost);
  } else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_edges_cost");
}

int grr_edges_compact(int64_t E, int32_t Nq, const uint32_t *mask, const int64_t *offsets, int32_t *out, void *stream) {
  if (E <= 0) return 0;
  if (!mask || !offsets || !out) { set_error("grr_edges_compact: null buffer"); return -1; }
  k_edges_compact<<<(unsigned)E, 128, 0, (cudaStream_t)stream>>>(Nq, mask, (const long long *)offsets, out);
  return launch_ok("k_edges_compact");
}

int grr_valid_edge_batch(grr_handle_t h, int dtype, int64_t P, const void *qa, const void *qb, const void *wa,
                         const void *wb, double epsilon, uint8_t *valid, int32_t *info, uint64_t *counters, void *stream) {
  GRR_CHECK_HANDLE(h);
  const Ops *ops = pick_ops(dtype, h->hc.mode);
  if (!ops) { set_error("bad dtype %d", dtype); return -1; }
  if (P > 0 && (!qa || !qb || !wa || !wb || !valid)) { set_error("grr_valid_edge_batch: null buffer"); return -1; }
  if (P <= 0) return 0;
  unsigned long long *work = work_slot(h);
  if (!work) return -3;
  return ops->valid_edge_batch(h->hc, P, qa, qb, wa, wb, epsilon, valid, info, work, (unsigned long long *)counters,
                               (cudaStream_t)stream);
}

int grr_gather_node_blocks(int dtype, int64_t K, const int32_t *idx, int32_t n, int32_t Nqp, const void *q,
                           const int32_t *count, void *packed, int32_t *packed_count, void *stream) {
  if (K <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GRR_F32)
    k_node_blocks<float, true><<<(unsigned)K, 128, 0, st>>>(K, idx, n * Nqp, (float *)q, (int32_t *)count, (float *)packed, packed_count);
  else if (dtype == GRR_F64)
    k_node_blocks<double, true><<<(unsigned)K, 128, 0, st>>>(K, idx, n * Nqp, (double *)q, (int32_t *)count, (double *)packed, packed_count);
  else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_node_blocks<gather>");
}

int grr_scatter_node_blocks(int dtype, int64_t K, const int32_t *idx, int32_t n, int32_t Nqp, const void *packed,
                            const int32_t *packed_count, void *q, int32_t *count, void *stream) {
  if (K <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == GRR_F32)
    k_node_blocks<float, false><<<(unsigned)K, 128, 0, st>>>(K, idx, n * Nqp, (float *)q, count, (float *)packed, (int32_t *)packed_count);
  else if (dtype == GRR_F64)
    k_node_blocks<double, false><<<(unsigned)K, 128, 0, st>>>(K, idx, n * Nqp, (double *)q, count, (double *)packed, (int32_t *)packed_count);
  else { set_error("bad dtype %d", dtype); return -1; }
  return launch_ok("k_node_blocks<scatter>");
}

// ------------------------------------------------------------------------------------------
// halo exchange over NCCL for hosts without PyTorch: the library resolves
// ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd from the libnccl.so.2 the process already
// uses (dlopen by soname), so it has no link-time dependency on NCCL.
// ------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
  bool ok;
};
const NcclApi *nccl_api() {
  static NcclApi api = {nullptr, nullptr, nullptr, nullptr, nullptr, false};
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  const char *path = getenv("GRR_NCCL_LIB");
  void *lib = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_error("grr_halo_exchange: cannot load %s: %s", path ? path : "libnccl.so.2", dlerror()); return nullptr; }
  api.GroupStart = (ncclResult_t(*)())dlsym(lib, "ncclGroupStart");
  api.GroupEnd = (ncclResult_t(*)())dlsym(lib, "ncclGroupEnd");
  api.Send = (ncclResult_t(*)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclSend");
  api.Recv = (ncclResult_t(*)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclRecv");
  api.GetErrorString = (const char *(*)(ncclResult_t))dlsym(lib, "ncclGetErrorString");
  api.ok = api.GroupStart && api.GroupEnd && api.Send && api.Recv;
  if (!api.ok) { set_error("grr_halo_exchange: libnccl lacks ncclSend / ncclRecv (NCCL >= 2.7 needed)"); return nullptr; }
  return &api;
}
}  // namespace

extern "C" int grr_halo_exchange(grr_handle_t h, void *nccl_comm, int dtype, int32_t Nqp, int32_t n_slabs, const int32_t *peer,
                                 const int32_t *is_send, const int64_t *node_lo, const int64_t *node_hi, void *q, int32_t *count,
                                 void *stream) {
  GRR_CHECK_HANDLE(h);
  if (n_slabs < 0 || Nqp < 1) { set_error("grr_halo_exchange: bad sizes"); return -1; }
  if (n_slabs == 0) return 0;
  if (!nccl_comm || !peer || !is_send || !node_lo || !node_hi || !q) { set_error("grr_halo_exchange: null argument"); return -1; }
  if (dtype != GRR_F32 && dtype != GRR_F64) { set_error("bad dtype %d", dtype); return -1; }
  for (int k = 0; k < n_slabs; k++)
    if (node_lo[k] < 0 || node_hi[k] < node_lo[k] || peer[k] < 0) { set_error("grr_halo_exchange: bad slab %d", k); return -1; }
  const NcclApi *api = nccl_api();
  if (!api) return -4;
  const size_t blk = (size_t)h->hc.n * (size_t)Nqp, esz = dtype == GRR_F64 ? 8 : 4;
  const ncclDataType_t ty = dtype == GRR_F64 ? ncclFloat64 : ncclFloat32;
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  cudaStream_t st = (cudaStream_t)stream;
  ncclResult_t rc = api->GroupStart();
  for (int k = 0; k < n_slabs && rc == ncclSuccess; k++) {
    const size_t nn = (size_t)(node_hi[k] - node_lo[k]);
    if (nn == 0) continue;
    char *qp = (char *)q + (size_t)node_lo[k] * blk * esz;
    rc = is_send[k] ? api->Send(qp, nn * blk, ty, peer[k], comm, st) : api->Recv(qp, nn * blk, ty, peer[k], comm, st);
    if (rc == ncclSuccess && count)
      rc = is_send[k] ? api->Send(count + node_lo[k], nn, ncclInt32, peer[k], comm, st)
                      : api->Recv(count + node_lo[k], nn, ncclInt32, peer[k], comm, st);
  }
  const ncclResult_t rc2 = api->GroupEnd();
  if (rc == ncclSuccess) rc = rc2;
  if (rc != ncclSuccess) { set_error("grr_halo_exchange: NCCL error %d (%s)", (int)rc, api->GetErrorString ? api->GetErrorString(rc) : "?"); return -4; }
  return 0;
}

static int fp32_peak_impl(int device, double *tflops_out, bool packed);
int grr_fp64_peak(int device, double *tflops_out) {
  if (!tflops_out) { set_error("null argument"); return -1; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); return -2; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 1024;
  double *buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) { set_error("cudaMalloc failed"); return -3; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    k_dfma_peak<<<blocks, threads>>>(buf, iters, 0.999, 0.001);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double tf = 2.0 * 8 * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(buf);
  if (cudaGetLastError() != cudaSuccess) { set_error("k_dfma_peak failed"); return -2; }
  *tflops_out = best;
  return 0;
}
int grr_fp32_peak(int device, double *tflops_out) { return fp32_peak_impl(device, tflops_out, false); }
int grr_fp32x2_peak(int device, double *tflops_out) { return fp32_peak_impl(device, tflops_out, true); }
}  // extern "C"

static int fp32_peak_impl(int device, double *tflops_out, bool packed) {
  if (!tflops_out) { set_error("null argument"); return -1; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); return -2; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float *buf = nullptr;
  if (cudaMalloc(&buf, sizeof(float) * blocks * threads) != cudaSuccess) { set_error("cudaMalloc failed"); return -3; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    if (packed) k_fma2_peak<<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
    else k_fma_peak<<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads * (packed ? 2.0 : 1.0);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(buf);
  int rc = launch_ok("k_fma_peak");
  *tflops_out = best;
  return rc;
}


