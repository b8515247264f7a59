// grr_vis.cu -- front end of the visibility-graph variant: everything around the edge predicate.
//
//   constructSelfMotionManifold  grr/solver.py:623-669   k_manifold       (one warp per workspace node)
//   testAndConnectQccs           grr/solver.py:517-561   k_qcc_candidates (one warp per CC pair of a workspace edge),
//   parallelEdgeConnection       grr/solver.py:949-982   k_qcc_wave, k_qcc_connect
//
// The reference interleaves two things: choosing which pairs to test (distance sorts, union-find, early exits) and
// validEdge itself.  validEdge is a pure function of the pair, so here the predicate is evaluated by the edge kernels
// on batches (all same-node pairs i > j through grr_edges_build on the workspace "edges" (i, i); explicit lists through
// grr_edges_test_listed) and these kernels replay the reference's sequential choice logic on the verdict bits.  The
// replay tests nothing itself; it reports the number of validEdge calls the reference's loop would have made.
//
// Distances are robot.distance in IEEE double without contraction (sum of squares in joint order, sqrt), ties in
// the reference's tuple order (dist, a, b), so that the orders equal those of Python's sorted().
#include <stdint.h>
#include <float.h>
#include <cuda_runtime.h>
#include "grr_device.cuh"
#include "grr_host.h"

namespace grr {

static inline int vis_launch_ok(const char *what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -2; }
  return 0;
}

template <typename Real>
__device__ __forceinline__ double q_distance(int n, unsigned spin, const Real *a, int sa, const Real *b, int sb) {
  double s = 0;
  if (spin == 0) {
    for (int j = 0; j < n; j++) {
      const double d = __dsub_rn((double)a[(long long)j * sa], (double)b[(long long)j * sb]);
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
  } else {
    // 'spin' joints: wrapped angular difference (robot.distance; same three functions as the oracle's restatement).  A loop of
    // its own: with the test inside the common loop the candidate kernel of chains without spin joints was 50 % slower.
    for (int j = 0; j < n; j++) {
      const double x = (double)a[(long long)j * sa], y = (double)b[(long long)j * sb];
      const double d = (spin >> j & 1u) ? angle_diff(x, y) : __dsub_rn(x, y);
      s = __dadd_rn(s, __dmul_rn(d, d));
    }
  }
  return __dsqrt_rn(s);
}

__device__ __forceinline__ bool key_less(double d1, unsigned j1, double d2, unsigned j2) {
  return d1 < d2 || (d1 == d2 && j1 < j2);
}
constexpr unsigned kNone = 0xffffffffu;

// minimum (d, j) over the warp, j == kNone <=> the lane has no candidate; all lanes get the result
__device__ __forceinline__ void warp_key_min(double &d, unsigned &j) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, o);
    const unsigned oj = __shfl_xor_sync(0xffffffffu, j, o);
    if (oj != kNone && (j == kNone || key_less(od, oj, d, j))) { d = od; j = oj; }
  }
}

// keyed hash standing in for Python's random.sample (solver.py:551, :973): the sampled candidates of a CC pair are the
// ones with the smallest hash (lowbias32 mixing of seed, workspace node ids and the pair)
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t pair_hash(uint32_t seed, uint32_t iw, uint32_t jw, uint32_t ab) {
  uint32_t h = mix32(seed ^ (iw * 0x9e3779b1u));
  h = mix32(h ^ (jw * 0x85ebca77u));
  return mix32(h ^ (ab * 0xc2b2ae3du));
}

// ------------------------------------------------------------------------------------------
// constructSelfMotionManifold (solver.py:623-669), one warp per workspace node.
//   knn == 0: "visibility prm with all-pairs connections" (:645-655): for i ascending, the j < i of other components in
//             increasing (distance, j); each is tested unless it joined i's component meanwhile.
//   knn  > 0: (:635-643): for i ascending, the knn nearest configs of the node in increasing (distance, j) (i itself
//             included, as KDTree.knearest returns it); tested when in another component.
// The verdict of validEdge(q_i, q_j, x, x) is bit (i, j) of the node's self mask.
//   list != nullptr: emit the pairs the knn mode may test instead ({node << 32 | i << 16 | j, 0}) and stop.
// Outputs: label[node][j] = index of j's component in qccs order (components by smallest member), ncc, parents
// (predecessor in a depth-first traversal of the merge forest from roots in index order, neighbours in the order the
// merges happened -- nx.dfs_edges(Gccs), :663-666), members[node][.] = configs sorted by (component, id),
// cc_start[node][c] = first position of component c in members (Nq + 1 entries), ntests = validEdge calls of the
// reference's loop.
// ------------------------------------------------------------------------------------------
template <typename Real, int R>
__global__ void __launch_bounds__(32)
k_manifold(int n, unsigned spin, int Nq, int Nqp, const Real *__restrict__ Q, const int32_t *__restrict__ count,
           const int32_t *__restrict__ node_list, const uint32_t *__restrict__ smask, int knn, int32_t *label_out,
           int32_t *ncc_out, int32_t *parents_out, int32_t *members_out, int32_t *cc_start_out, int32_t *ntests_out,
           unsigned long long *list, unsigned long long cap) {
  extern __shared__ int32_t sm[];
  int32_t *label = sm, *medge = sm + Nq, *par = sm + 2 * Nq, *seen = sm + 3 * Nq, *stv = sm + 4 * Nq, *stp = sm + 5 * Nq;
  const int node = node_list ? node_list[blockIdx.x] : (int)blockIdx.x;
  const int lane = threadIdx.x;
  const int c = count[node];
  const int Wd = (Nq + 31) >> 5;
  const Real *q = Q + (long long)node * n * Nqp;
  const uint32_t *mrow = smask ? smask + (long long)node * Nq * Wd : nullptr;
  for (int j = lane; j < Nq; j += 32) { label[j] = j; par[j] = -1; seen[j] = 0; }
  __syncwarp();
  int nm = 0, tests = 0;
  for (int i = (knn > 0 ? 0 : 1); i < c; i++) {
    double d[R];
    bool alive[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int j = lane + 32 * r;
      alive[r] = knn > 0 ? (j < c) : (j < i);
      d[r] = alive[r] ? q_distance(n, spin, q + i, Nqp, q + j, Nqp) : 0.0;
    }
    const int rounds = knn > 0 ? min(knn, c) : i;
    for (int round = 0; round < rounds; round++) {
      const int li = label[i];
      double bd = DBL_MAX;
      unsigned bj = kNone;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const unsigned j = lane + 32 * r;
        if (!alive[r]) continue;
        if (knn == 0 && label[j] == li) { alive[r] = false; continue; }           // :652 (joined i's component meanwhile)
        if (bj == kNone || key_less(d[r], j, bd, bj)) { bd = d[r]; bj = j; }
      }
      warp_key_min(bd, bj);
      if (bj == kNone) break;
#pragma unroll
      for (int r = 0; r < R; r++) if (lane + 32 * r == (int)bj) alive[r] = false;
      if (list) {
        if (lane == 0 && (int)bj != i) {
          const unsigned long long pos = atomicAdd(list, 1ull);
          if (pos < cap) {
            list[2 + 2 * pos] = ((unsigned long long)(unsigned)node << 32) | ((unsigned long long)(unsigned)i << 16) | (unsigned long long)bj;
            list[3 + 2 * pos] = 0ull;
          }
        }
        continue;
      }
      GRR_DCHECK((int)bj < c && i < c);
      const int lj = label[bj];
      if (lj == li) continue;                                                      // :639 (knn mode)
      tests++;
      const bool valid = (mrow[i * Wd + (bj >> 5)] >> (bj & 31)) & 1u;            // validEdge(q_i, q_j, x_i, x_i)
      if (valid) {
        const int keep = min(li, lj), drop = max(li, lj);
        __syncwarp();
        for (int x = lane; x < c; x += 32) if (label[x] == drop) label[x] = keep;
        GRR_DCHECK(nm < Nq - 1 || Nq == 1);
        if (lane == 0) medge[nm] = (i << 16) | (int)bj;
        nm++;
        __syncwarp();
      }
    }
  }
  if (list) return;
  __syncwarp();
  // components in qccs order: by smallest member (= the label, kept minimal by the merges)
  int ncc = 0;
  for (int j = 0; j < c; j++) ncc += (label[j] == j);
  int32_t *lab_o = label_out + (long long)node * Nq, *mem_o = members_out + (long long)node * Nq;
  int32_t *ccs_o = cc_start_out + (long long)node * (Nq + 1);
  for (int j = lane; j < Nq; j += 32) {
    int idx = -1;
    if (j < c) { idx = 0; const int lj = label[j]; for (int x = 0; x < lj; x++) idx += (label[x] == x); }
    stv[j] = idx;                                   // component index of j
  }
  __syncwarp();
  for (int j = lane; j < Nq; j += 32) {
    lab_o[j] = stv[j];
    if (j < c) {
      int pos = 0;
      const int cj = stv[j];
      for (int x = 0; x < c; x++) pos += (stv[x] < cj) || (stv[x] == cj && x < j);
      mem_o[pos] = j;
    } else mem_o[j] = -1;
  }
  for (int cc = lane; cc <= Nq; cc += 32) {
    int pos = 0;
    for (int x = 0; x < c; x++) pos += (stv[x] < cc);
    ccs_o[cc] = pos;
  }
  __syncwarp();
  // depth-first traversal of the merge forest (lane 0): roots in index order, neighbours in merge order
  if (lane == 0) {
    for (int root = 0; root < c; root++) {
      if (seen[root]) continue;
      seen[root] = 1;
      int sp = 0;
      stv[0] = root; stp[0] = 0;
      while (sp >= 0) {
        const int v = stv[sp];
        bool found = false;
        for (int m = stp[sp]; m < nm; m++) {
          const int a = medge[m] >> 16, b = medge[m] & 0xffff;
          const int w = (a == v) ? b : ((b == v) ? a : -1);
          if (w >= 0 && !seen[w]) {
            stp[sp] = m + 1; seen[w] = 1; par[w] = v;
            sp++; stv[sp] = w; stp[sp] = 0;
            found = true;
            break;
          }
        }
        if (!found) sp--;
      }
    }
    ncc_out[node] = ncc;
    if (ntests_out) ntests_out[node] = tests;
  }
  __syncwarp();
  int32_t *par_o = parents_out + (long long)node * Nq;
  for (int j = lane; j < Nq; j += 32) par_o[j] = par[j];
}

// ------------------------------------------------------------------------------------------
// Candidates of every CC pair of a workspace edge (solver.py:529-552, :954-974), one warp per CC pair:
// slots [0, T): the pairs (a in qicc, b in qjcc) with the T smallest (distance, a, b); slots [T, 2T): of the
// remaining pairs the T with the smallest (pair_hash, a, b), in that order (the reference draws them with random.sample).
// cand[(cc_off[e] + ci * ncc[jw] + cj) * 2T + slot] = a << 16 | b, or -1.
// ------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(128)
k_qcc_candidates(long long E, const int32_t *__restrict__ wedges, int n, unsigned spin, int Nq, int Nqp, const Real *__restrict__ Q,
                 const int32_t *__restrict__ uid, const int32_t *__restrict__ ncc, const int32_t *__restrict__ members,
                 const int32_t *__restrict__ cc_start, const long long *__restrict__ cc_off, int T, uint32_t seed,
                 int swap, int32_t *__restrict__ cand) {
  const long long e = blockIdx.x;
  const int iw = wedges[2 * e + (swap ? 1 : 0)], jw = wedges[2 * e + (swap ? 0 : 1)];     // outer-loop node, inner-loop node
  const int ni = ncc[iw], nj = ncc[jw];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const Real *qi = Q + (long long)iw * n * Nqp, *qj = Q + (long long)jw * n * Nqp;
  const int32_t *mi = members + (long long)iw * Nq, *mj = members + (long long)jw * Nq;
  const int32_t *si = cc_start + (long long)iw * (Nq + 1), *sj = cc_start + (long long)jw * (Nq + 1);
  const uint32_t ui = uid ? (uint32_t)uid[iw] : (uint32_t)iw, uj = uid ? (uint32_t)uid[jw] : (uint32_t)jw;
  for (int p = warp; p < ni * nj; p += nwarps) {
    const int ci = p / nj, cj = p - ci * nj;
    const int a0 = si[ci], na = si[ci + 1] - a0, b0 = sj[cj], nb = sj[cj + 1] - b0;
    const int total = na * nb;
    int32_t *out = cand + (cc_off[e] + p) * 2ll * T;
    // phase 1: T smallest (distance, a, b)
    double last_d = -1.0;
    unsigned last_k = 0;
    bool have_last = false;
    int got = 0;
    for (; got < T && got < total; got++) {
      double bd = DBL_MAX;
      unsigned bk = kNone;
      for (int t = lane; t < total; t += 32) {
        const int a = mi[a0 + t / nb], b = mj[b0 + t % nb];
        const unsigned key = ((unsigned)a << 16) | (unsigned)b;
        const double d = q_distance(n, spin, qi + a, Nqp, qj + b, Nqp);
        if (have_last && !key_less(last_d, last_k, d, key)) continue;          // already selected
        if (bk == kNone || key_less(d, key, bd, bk)) { bd = d; bk = key; }
      }
      warp_key_min(bd, bk);
      GRR_DCHECK(bk != kNone && (int)(bk >> 16) < Nq && (int)(bk & 0xffffu) < Nq);
      if (lane == 0) out[got] = (int32_t)bk;
      last_d = bd; last_k = bk; have_last = true;
    }
    for (int s = got + lane; s < T; s += 32) out[s] = -1;
    // phase 2: of the rest, T smallest (hash, a, b)
    int got2 = 0;
    if (total > T) {
      unsigned last_h = 0, last_hk = 0;
      bool have_h = false;
      for (; got2 < T && got2 < total - T; got2++) {
        unsigned bh = 0, bk = kNone;
        for (int t = lane; t < total; t += 32) {
          const int a = mi[a0 + t / nb], b = mj[b0 + t % nb];
          const unsigned key = ((unsigned)a << 16) | (unsigned)b;
          const double d = q_distance(n, spin, qi + a, Nqp, qj + b, Nqp);
          if (!key_less(last_d, last_k, d, key)) continue;                       // one of the T nearest
          const unsigned h = pair_hash(seed, ui, uj, key);
          if (have_h && !(h > last_h || (h == last_h && key > last_hk))) continue;
          if (bk == kNone || h < bh || (h == bh && key < bk)) { bh = h; bk = key; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned oh = __shfl_xor_sync(0xffffffffu, bh, o), ok = __shfl_xor_sync(0xffffffffu, bk, o);
          if (ok != kNone && (bk == kNone || oh < bh || (oh == bh && ok < bk))) { bh = oh; bk = ok; }
        }
        if (lane == 0) out[T + got2] = (int32_t)bk;
        last_h = bh; last_hk = bk; have_h = true;
      }
    }
    for (int s = T + got2 + lane; s < 2 * T; s += 32) out[s] = -1;
  }
}

// candidate (a of the outer-loop node) << 16 | (b of the inner-loop node) -> (config of wedges[e][0]) << 16 | (config of wedges[e][1])
__device__ __forceinline__ int32_t canon(int32_t ab, int swap) { return swap ? (int32_t)(((uint32_t)ab << 16) | ((uint32_t)ab >> 16)) : ab; }

__device__ __forceinline__ bool vbit(const uint32_t *vm, long long e, int Nq, int Wd, int32_t ab) {
  const int a = ab >> 16, b = ab & 0xffff;
  GRR_DCHECK(a >= 0 && a < Nq && b < Nq);
  return (vm[(e * Nq + a) * Wd + (b >> 5)] >> (b & 31)) & 1u;
}

// One wave of tests: for every CC pair none of whose candidate slots [0, s0) was valid, list slots [s0, s1).
__global__ void __launch_bounds__(128)
k_qcc_wave(long long E, const int32_t *__restrict__ wedges, const int32_t *__restrict__ ncc,
           const long long *__restrict__ cc_off, int T2, const int32_t *__restrict__ cand, int Nq,
           const uint32_t *__restrict__ vmask, int s0, int s1, int swap, unsigned long long *list, unsigned long long cap) {
  const long long e = blockIdx.x;
  const int Wd = (Nq + 31) >> 5;
  const int np = ncc[wedges[2 * e]] * ncc[wedges[2 * e + 1]];
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    const int32_t *c = cand + (cc_off[e] + p) * (long long)T2;
    bool hit = false;
    for (int s = 0; s < s0 && !hit; s++) if (c[s] >= 0 && vbit(vmask, e, Nq, Wd, canon(c[s], swap))) hit = true;
    if (hit) continue;
    for (int s = s0; s < s1; s++) {
      if (c[s] < 0) continue;
      const unsigned long long pos = atomicAdd(list, 1ull);
      if (pos < cap) {
        list[2 + 2 * pos] = ((unsigned long long)e << 32) | (unsigned long long)(uint32_t)canon(c[s], swap);
        list[3 + 2 * pos] = 0ull;
      }
    }
  }
}

// Replay of the CC-pair loops on the verdict bits, one thread per workspace edge.
//   mode 0 = testAndConnectQccs (:530-559): a component of node i stops at the first component of node j that one of the
//            nearest candidates connects it to (a connection by a sampled candidate does not end the loop, :551-558);
//   mode 1 = parallelEdgeConnection (:953-980): every CC pair is tried.
// conn[cc_off[e] + p] = the connecting a << 16 | b or -1 (-2: not reached), bit (a, b) set in mask; nconn[e] =
// numConnected; ntests[e] = validEdge calls of the reference's loops.
__global__ void __launch_bounds__(128)
k_qcc_connect(long long E, const int32_t *__restrict__ wedges, const int32_t *__restrict__ ncc,
              const long long *__restrict__ cc_off, int T, const int32_t *__restrict__ cand, int Nq,
              const uint32_t *__restrict__ vmask, int mode, int swap, uint32_t *mask, int32_t *__restrict__ conn,
              int32_t *__restrict__ nconn, int32_t *__restrict__ ntests) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int Wd = (Nq + 31) >> 5;
  const int ni = ncc[wedges[2 * e + (swap ? 1 : 0)]], nj = ncc[wedges[2 * e + (swap ? 0 : 1)]];
  int connected = 0, tests = 0;
  for (int ci = 0; ci < ni; ci++) {
    bool done = false;
    for (int cj = 0; cj < nj; cj++) {
      const long long p = cc_off[e] + (long long)ci * nj + cj;
      if (done) { conn[p] = -2; continue; }
      const int32_t *c = cand + p * 2ll * T;
      int32_t hit = -1;
      int hit_slot = 0;
      for (int s = 0; s < 2 * T && hit < 0; s++) {
        if (c[s] < 0) continue;
        tests++;
        if (vbit(vmask, e, Nq, Wd, canon(c[s], swap))) { hit = c[s]; hit_slot = s; }
      }
      conn[p] = hit;
      if (hit >= 0) {
        const int32_t hc = canon(hit, swap);
        const int a = hc >> 16, b = hc & 0xffff;
        atomicOr(&mask[(e * Nq + a) * Wd + (b >> 5)], 1u << (b & 31));
        connected++;
        if (mode == 0 && hit_slot < T) done = true;                              // :548-550 (the sampled phase has no break, :551-558)
      }
    }
  }
  nconn[e] = connected;
  if (ntests) ntests[e] = tests;
}

template <typename Real>
static int manifold_launch(int64_t Nn, const int32_t *node_list, int n, unsigned spin, int Nq, int Nqp, const void *Q, const int32_t *count,
                           const uint32_t *smask, int knn, int32_t *label, int32_t *ncc, int32_t *parents, int32_t *members,
                           int32_t *cc_start, int32_t *ntests, unsigned long long *list, unsigned long long cap, cudaStream_t st) {
  const size_t smem = 6 * (size_t)Nq * sizeof(int32_t);
  auto launch = [&](auto kern) -> int {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("k_manifold: Nq=%d does not fit shared memory", Nq); return -2;
    }
    kern<<<(unsigned)Nn, 32, smem, st>>>(n, spin, Nq, Nqp, (const Real *)Q, count, node_list, smask, knn, label, ncc, parents, members,
                                         cc_start, ntests, list, cap);
    return vis_launch_ok("k_manifold");
  };
  if (Nq <= 128) return launch(k_manifold<Real, 4>);
  if (Nq <= 256) return launch(k_manifold<Real, 8>);
  if (Nq <= 512) return launch(k_manifold<Real, 16>);
  set_error("k_manifold: Nq=%d > 512 configs per node", Nq);
  return -1;
}

}  // namespace grr

using namespace grr;

extern "C" {

int grr_manifold_knn_list(int dtype, int64_t Nn, const int32_t *node_list, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                          const int32_t *count, int32_t knn, uint64_t *list, uint64_t cap, uint32_t spin_mask, void *stream) {
  if (Nn <= 0) return 0;
  if (!q_kept || !count || !list || knn < 1 || n < 1 || Nq < 1 || Nqp < Nq || Nq > 65535) { set_error("grr_manifold_knn_list: bad argument"); return -1; }
  if (dtype == GRR_F32)
    return manifold_launch<float>(Nn, node_list, n, spin_mask, Nq, Nqp, q_kept, count, nullptr, knn, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, (unsigned long long *)list, cap, (cudaStream_t)stream);
  if (dtype == GRR_F64)
    return manifold_launch<double>(Nn, node_list, n, spin_mask, Nq, Nqp, q_kept, count, nullptr, knn, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   nullptr, (unsigned long long *)list, cap, (cudaStream_t)stream);
  set_error("bad dtype %d", dtype);
  return -1;
}

int grr_manifold_components(int dtype, int64_t Nn, const int32_t *node_list, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                            const int32_t *count, const uint32_t *self_mask, int32_t knn, int32_t *label, int32_t *ncc,
                            int32_t *parents, int32_t *members, int32_t *cc_start, int32_t *ntests, uint32_t spin_mask, void *stream) {
  if (Nn <= 0) return 0;
  if (!q_kept || !count || !self_mask || !label || !ncc || !parents || !members || !cc_start || knn < 0 || n < 1 || Nq < 1 ||
      Nqp < Nq || Nq > 65535) { set_error("grr_manifold_components: bad argument"); return -1; }
  if (dtype == GRR_F32)
    return manifold_launch<float>(Nn, node_list, n, spin_mask, Nq, Nqp, q_kept, count, self_mask, knn, label, ncc, parents, members, cc_start,
                                  ntests, nullptr, 0, (cudaStream_t)stream);
  if (dtype == GRR_F64)
    return manifold_launch<double>(Nn, node_list, n, spin_mask, Nq, Nqp, q_kept, count, self_mask, knn, label, ncc, parents, members, cc_start,
                                   ntests, nullptr, 0, (cudaStream_t)stream);
  set_error("bad dtype %d", dtype);
  return -1;
}

int grr_qcc_candidates(int dtype, int64_t E, const int32_t *wedges, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                       const int32_t *node_uid, const int32_t *ncc, const int32_t *members, const int32_t *cc_start,
                       const int64_t *cc_off, int32_t max_tests, uint64_t seed, int32_t swap, int32_t *cand, uint32_t spin_mask, void *stream) {
  if (E <= 0) return 0;
  if (!wedges || !q_kept || !ncc || !members || !cc_start || !cc_off || !cand || max_tests < 1 || Nq > 65535) {
    set_error("grr_qcc_candidates: bad argument"); return -1;
  }
  const uint32_t s32 = (uint32_t)seed ^ (uint32_t)(seed >> 32);
  if (dtype == GRR_F32)
    k_qcc_candidates<float><<<(unsigned)E, 128, 0, (cudaStream_t)stream>>>(E, wedges, n, spin_mask, Nq, Nqp, (const float *)q_kept, node_uid, ncc,
                                                                          members, cc_start, (const long long *)cc_off, max_tests, s32, swap, cand);
  else if (dtype == GRR_F64)
    k_qcc_candidates<double><<<(unsigned)E, 128, 0, (cudaStream_t)stream>>>(E, wedges, n, spin_mask, Nq, Nqp, (const double *)q_kept, node_uid, ncc,
                                                                           members, cc_start, (const long long *)cc_off, max_tests, s32, swap, cand);
  else { set_error("bad dtype %d", dtype); return -1; }
  return vis_launch_ok("k_qcc_candidates");
}

int grr_qcc_wave(int64_t E, const int32_t *wedges, const int32_t *ncc, const int64_t *cc_off, int32_t max_tests, const int32_t *cand,
                 int32_t Nq, const uint32_t *tested_mask, int32_t slot0, int32_t slot1, int32_t swap, uint64_t *list, uint64_t cap,
                 void *stream) {
  if (E <= 0) return 0;
  if (!wedges || !ncc || !cc_off || !cand || !tested_mask || !list || slot0 < 0 || slot1 > 2 * max_tests || slot0 > slot1) {
    set_error("grr_qcc_wave: bad argument"); return -1;
  }
  k_qcc_wave<<<(unsigned)E, 128, 0, (cudaStream_t)stream>>>(E, wedges, ncc, (const long long *)cc_off, 2 * max_tests, cand, Nq, tested_mask,
                                                            slot0, slot1, swap, (unsigned long long *)list, cap);
  return vis_launch_ok("k_qcc_wave");
}

int grr_qcc_connect(int64_t E, const int32_t *wedges, const int32_t *ncc, const int64_t *cc_off, int32_t max_tests, const int32_t *cand,
                    int32_t Nq, const uint32_t *tested_mask, int32_t mode, int32_t swap, uint32_t *mask, int32_t *conn,
                    int32_t *nconn, int32_t *ntests, void *stream) {
  if (E <= 0) return 0;
  if (!wedges || !ncc || !cc_off || !cand || !tested_mask || !mask || !conn || !nconn || (mode != 0 && mode != 1)) {
    set_error("grr_qcc_connect: bad argument"); return -1;
  }
  k_qcc_connect<<<(unsigned)((E + 127) / 128), 128, 0, (cudaStream_t)stream>>>(E, wedges, ncc, (const long long *)cc_off, max_tests, cand, Nq,
                                                                               tested_mask, mode, swap, mask, conn, nconn, ntests);
  return vis_launch_ok("k_qcc_connect");
}

uint32_t grr_pair_hash(uint64_t seed, uint32_t iw, uint32_t jw, uint32_t a, uint32_t b) {
  return pair_hash((uint32_t)seed ^ (uint32_t)(seed >> 32), iw, jw, (a << 16) | b);
}

}  // extern "C"
