// grr_csp.cu -- the roadmap's immediate consumer on the device bitmasks: the constraint
// satisfaction problem of RedundancyResolver.makeCSP (grr/solver.py:984-1022: one variable per workspace
// node with configs, domain = config indices; one table constraint per workspace edge whose C-space edge set
// is non-empty, allowed tuples = that set) and the min-conflicts descent of
// CSP.randomDescentAssignment (grr/utils/csp.py:934-1041) with CSP.numIncompatible (csp.py:133-150) and
// CSP.numConflicts (csp.py:79-84).
//
// The table of a constraint IS the Nq x Nq bitmask grr_edges_build wrote: (iq, jq) allowed <=> bit jq of row
// iq.  numIncompatible(var, value) is therefore one bit test per incident workspace edge, and the descent
// evaluates all values of a node at once, one value per lane.
//
// The reference visits the conflicting variables one by one in a shuffled order and updates conflict counts
// immediately (Gauss-Seidel).  Variables that are not adjacent in the workspace graph do not see each
// other, so all variables of one colour class of a proper colouring can be processed together with exactly
// the result of visiting them one after the other: one sweep = the reference's pass with the order
// "colour 0 nodes, colour 1 nodes, ...".  The reference's `active` set is kept as it is there (csp.py:978,
// :993-995, :1019-1020, :1027-1028): a pass visits the variables that were active when it began; a variable
// leaves the set when visited and enters it when a neighbour moves.  G independent starting guesses are
// descended at once (solveCSP's numInitialGuesses, solver.py:2048-2056).
#include "grr_device.cuh"
#include "grr_host.h"

namespace grr {

static inline int csp_launch_ok(const char *what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -2; }
  return 0;
}

// conflict of (node = val) with neighbour entry k: adj_edge[k] = e if node is the first endpoint of workspace
// edge e (row index = this node's value), ~e if it is the second (column index)
__device__ __forceinline__ int csp_conflict(const uint32_t *__restrict__ mask, const uint8_t *__restrict__ edge_active,
                                            int Nq, int W, int ee, int val, int other) {
  const bool second = ee < 0;
  const long long e = second ? ~ee : ee;
  if (!edge_active[e] || other < 0) return 0;           // no constraint (solver.py:1019) / unassigned (csp.py:144)
  const int r = second ? other : val, cbit = second ? val : other;
  GRR_DCHECK(r >= 0 && r < Nq && cbit >= 0 && cbit < Nq);
  const uint32_t w = mask[(e * Nq + r) * W + (cbit >> 5)];
  return ((w >> (cbit & 31)) & 1u) ? 0 : 1;
}

// random.choice(domain) for every variable (csp.py:950-955) from the counter-based generator of the
// sampling kernels: Philox4x32-10, key = seed, counter = (node, guess, 0, stream 2)
__global__ void k_csp_random(int Nw, int G, const int32_t *__restrict__ count, unsigned long long seed, int guess_offset,
                             int32_t *__restrict__ asg) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Nw * G) return;
  const int g = (int)(t / Nw), i = (int)(t - (long long)g * Nw);
  const int c = count[i];
  int v = -1;
  if (c > 0) {
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)(g + guess_offset), 0u, 2u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    v = (int)(r[0] % (uint32_t)c);
  }
  asg[t] = v;
}

// numIncompatible(node, assignment[node]) for every node (csp.py:133-150) and numConflicts (csp.py:79-84):
// every violated constraint is counted once, at its first endpoint.
__global__ void k_csp_conflicts(int Nw, int Nq, int W, int G, const int32_t *__restrict__ adj_ptr,
                                const int32_t *__restrict__ adj_node, const int32_t *__restrict__ adj_edge,
                                const uint32_t *__restrict__ mask, const uint8_t *__restrict__ edge_active,
                                const int32_t *__restrict__ asg, int32_t *__restrict__ node_conflicts,
                                int32_t *__restrict__ total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int mine = 0;
  if (t < (long long)Nw * G) {
    const int g = (int)(t / Nw), i = (int)(t - (long long)g * Nw);
    const int32_t *a = asg + (long long)g * Nw;
    const int val = a[i];
    int nc = 0;
    if (val >= 0) {
      for (int k = adj_ptr[i]; k < adj_ptr[i + 1]; k++) {
        const int cf = csp_conflict(mask, edge_active, Nq, W, adj_edge[k], val, a[adj_node[k]]);
        nc += cf;
        if (adj_edge[k] >= 0) mine += cf;
      }
    }
    if (node_conflicts) node_conflicts[t] = nc;
    if (total && mine) atomicAdd(&total[g], mine);
  }
}

// One descent visit of every node of `node_list` (one colour class) for every guess: one warp per
// (node, guess), one candidate value per lane.  New value = the first value with the fewest conflicts, taken
// only if it has strictly fewer conflicts than the current one (csp.py:1003-1012; assignmentCost is the
// base class's constant 0, csp.py:76-78).
__global__ void __launch_bounds__(128)
k_csp_sweep(int n_list, const int32_t *__restrict__ node_list, int Nw, int Nq, int W, int G,
            const int32_t *__restrict__ adj_ptr, const int32_t *__restrict__ adj_node,
            const int32_t *__restrict__ adj_edge, const uint32_t *__restrict__ mask,
            const uint8_t *__restrict__ edge_active, const int32_t *__restrict__ count, int32_t *asg,
            const uint8_t *__restrict__ in_order, uint8_t *active, int32_t *__restrict__ changed) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)n_list * G) return;
  const int g = (int)(w / n_list), i = node_list[w - (long long)g * n_list];
  GRR_DCHECK(i >= 0 && i < Nw && g < G);
  int32_t *a = asg + (long long)g * Nw;
  const int c = count[i];
  const int cur = a[i];
  if (c <= 0 || cur < 0 || !in_order[(long long)g * Nw + i]) return;
  uint8_t *act = active + (long long)g * Nw;
  const int k0 = adj_ptr[i], k1 = adj_ptr[i + 1];
  unsigned best = 0xffffffffu;                        // (conflicts << 16 | value): min = fewest, then first
  unsigned cur_nc = 0;
  for (int v0 = 0; v0 < c; v0 += 32) {
    const int val = v0 + lane;
    unsigned nc = 0xffffu;
    if (val < c) {
      nc = 0;
      for (int k = k0; k < k1; k++) nc += csp_conflict(mask, edge_active, Nq, W, adj_edge[k], val, a[adj_node[k]]);
      best = min(best, (nc << 16) | (unsigned)val);
    }
    const unsigned has = __ballot_sync(0xffffffffu, val == cur);
    if (has) cur_nc = __shfl_sync(0xffffffffu, nc, __ffs(has) - 1);
  }
  best = __reduce_min_sync(0xffffffffu, best);
  const bool move = (best >> 16) < cur_nc;
  if (move) {
    // wake the neighbours (csp.py:1014-1020); they are in other colour classes, i.e. not being visited now
    for (int k = k0 + lane; k < k1; k += 32) {
      const int ee = adj_edge[k];
      if (edge_active[ee < 0 ? ~ee : ee] && a[adj_node[k]] >= 0) act[adj_node[k]] = 1;
    }
  }
  __syncwarp();
  if (lane == 0) {
    if (move) { a[i] = (int)(best & 0xffffu); atomicAdd(&changed[g], 1); }
    act[i] = 0;                                       // csp.py:1027-1028
  }
}

}  // namespace grr

using namespace grr;

extern "C" {

int grr_csp_random_assignment(int32_t Nw, int32_t G, const int32_t *count, uint64_t seed, int32_t guess_offset,
                              int32_t *assignment, void *stream) {
  if (Nw <= 0 || G <= 0) return 0;
  if (!count || !assignment) { set_error("grr_csp_random_assignment: null buffer"); return -1; }
  const long long tot = (long long)Nw * G;
  k_csp_random<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Nw, G, count, seed, guess_offset, assignment);
  return csp_launch_ok("k_csp_random");
}

int grr_csp_conflicts(int32_t Nw, int32_t Nq, int32_t G, const int32_t *adj_ptr, const int32_t *adj_node,
                      const int32_t *adj_edge, const uint32_t *mask, const uint8_t *edge_active,
                      const int32_t *assignment, int32_t *node_conflicts, int32_t *total, void *stream) {
  if (Nw <= 0 || G <= 0) return 0;
  if (!adj_ptr || !adj_node || !adj_edge || !mask || !edge_active || !assignment) { set_error("grr_csp_conflicts: null buffer"); return -1; }
  if (Nq <= 0 || Nq > 65535) { set_error("grr_csp_conflicts: Nq=%d outside 1..65535", Nq); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (total && cudaMemsetAsync(total, 0, sizeof(int32_t) * (size_t)G, st) != cudaSuccess) { set_error("grr_csp_conflicts: memset failed"); return -2; }
  const long long tot = (long long)Nw * G;
  k_csp_conflicts<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(Nw, Nq, (Nq + 31) >> 5, G, adj_ptr, adj_node, adj_edge, mask,
                                                                   edge_active, assignment, node_conflicts, total);
  return csp_launch_ok("k_csp_conflicts");
}

int grr_csp_descent_sweep(int32_t n_list, const int32_t *node_list, int32_t Nw, int32_t Nq, int32_t G,
                          const int32_t *adj_ptr, const int32_t *adj_node, const int32_t *adj_edge, const uint32_t *mask,
                          const uint8_t *edge_active, const int32_t *count, int32_t *assignment,
                          const uint8_t *in_order, uint8_t *active, int32_t *changed, void *stream) {
  if (n_list <= 0 || G <= 0) return 0;
  if (!node_list || !adj_ptr || !adj_node || !adj_edge || !mask || !edge_active || !count || !assignment || !in_order ||
      !active || !changed) {
    set_error("grr_csp_descent_sweep: null buffer"); return -1;
  }
  if (Nq <= 0 || Nq > 65535) { set_error("grr_csp_descent_sweep: Nq=%d outside 1..65535", Nq); return -1; }
  const long long warps = (long long)n_list * G;
  k_csp_sweep<<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(n_list, node_list, Nw, Nq, (Nq + 31) >> 5, G, adj_ptr,
                                                                            adj_node, adj_edge, mask, edge_active, count,
                                                                            assignment, in_order, active, changed);
  return csp_launch_ok("k_csp_sweep");
}

}  // extern "C"
