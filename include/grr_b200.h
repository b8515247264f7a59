/*
 * grr_b200.h -- C ABI of libgrr_b200.so: the B200 (sm_100a) roadmap-construction hot path of
 * GlobalRedundancyResolution (per-workspace-node IK configuration sampling + the basic solver's
 * all-pairs C-space edge construction).
 *
 * The reference (krishauser/GlobalRedundancyResolution) has no FFI / plugin interface: the path
 * sits behind Python methods that call Klamp't/IKDB one instance at a time.  Each entry point
 * below states the reference interface (file:line under the reference root) it replaces.  The
 * Python classes in globalredundancyresolution_b200/{graph,solver}.py keep the reference's method
 * names and call these through ctypes (see INTEGRATION.md for the binding a maintainer would add
 * to the reference itself).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; grr_last_error() gives the message
 *     (thread-local).  Nothing throws.
 *   - all array arguments marked [dev] are caller-owned DEVICE pointers (e.g. torch tensors'
 *     data_ptr()); [host] are host pointers.  `stream` is a cudaStream_t (0 = default stream).
 *     Calls are asynchronous on `stream` unless stated otherwise.
 *   - a handle serves ONE host thread and ONE stream at a time (its ring of device work counters and its scratch
 *     buffer are not synchronised between callers): use one handle per concurrent stream, as the reference uses one
 *     robot / IK solver per worker process (solver.py:2379-2400).
 *   - `dtype` selects the arithmetic type of the path and of every "Real" buffer:
 *     GRR_F32 (the product path, FP32 SIMT) or GRR_F64 (verification build of the same kernels).
 *   - rotation matrices are ROW-major 3x3.
 *   - config storage ("node-blocked SoA"): Q[node][dof][Nqp], Nqp = Nq rounded up to a multiple
 *     of 4; a node's block is one contiguous n*Nqp chunk (one bulk copy into shared memory, one
 *     message in the halo exchange).  Batches of unrelated configs use plain SoA q[dof][P].
 *   - workspace parameters: params[node][dw], dw = 3 (free, fixed) or 6 (variable: position +
 *     rotation moment), as produced by sampleWorkspace (grr/graph.py:1110-1123,1151-1164).
 */
#ifndef GRR_B200_H_
#define GRR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRR_MAX_JOINTS 32

enum { GRR_F32 = 0, GRR_F64 = 1 };
enum { GRR_FREE = 0, GRR_FIXED = 1, GRR_VARIABLE = 2 };          /* IKTask orientation, graph.py:59-74 */
enum { GRR_IK_OK = 0, GRR_IK_STALL = 1, GRR_IK_CONVX = 2, GRR_IK_MAXIT = 3, GRR_IK_SKIPPED = 255 };

typedef struct grr_handle_s *grr_handle_t;

/* Serial kinematic chain from the robot base to the goal link with inactive links folded into
 * the constant transforms; replaces the Klamp't RobotModel + IKObjective + IKDB IKProblem /
 * IKSolverParams objects built in grr/graph.py:494-509,518-640. */
typedef struct {
  int32_t n;            /* active revolute joints, 1..GRR_MAX_JOINTS                              */
  int32_t mode;         /* GRR_FREE / GRR_FIXED / GRR_VARIABLE                                    */
  int32_t nlinks_eps;   /* "nlinks" of validEdgeLinear, graph.py:944                              */
  int32_t max_iters;    /* IKSolverParams.numIters                                                */
  double tol;           /* IKSolverParams.tol                                                     */
  double lambda;        /* Newton damping                                                         */
  const double *Rp;     /* [host] [n][9] constant rotation preceding joint j                       */
  const double *tp;     /* [host] [n][3] constant translation preceding joint j                    */
  const double *axis;   /* [host] [n][3] unit joint axis, local                                    */
  const double *qmin;   /* [host] [n]                                                              */
  const double *qmax;   /* [host] [n]                                                              */
  double ee_local[3];   /* constrained point in the last active joint's frame (graph.py:60,65,72)  */
  double R_tail[9];     /* goal-link frame relative to the last active joint's frame               */
  double R_goal[9];     /* fixed orientation goal (graph.py:71), ignored unless mode==GRR_FIXED    */
} grr_chain_desc;

/* Counters accumulated (atomically) by the kernels; layout of the `counters` arguments. */
enum {
  GRR_CNT_IK_SOLVES = 0,   /* IK instances run                                  */
  GRR_CNT_IK_ITERS = 1,    /* Newton iterations                                 */
  GRR_CNT_IK_EVALS = 2,    /* FK/residual(+Jacobian) passes                     */
  GRR_CNT_PAIRS = 3,       /* validEdge calls (pair tests)                      */
  GRR_CNT_VALID = 4,       /* pair tests that returned True                     */
  GRR_CNT_IK_OK = 5,       /* IK instances that converged                       */
  GRR_CNT_FLAGGED = 6,     /* grr_edges_build_exact: pairs re-tested in double precision              */
  GRR_CNT_FLIPPED = 7,     /* ... of which the double-precision verdict differed from the first pass  */
  GRR_CNT_FLAG_DROPPED = 8,/* ... near-tie pairs that did not fit the list (must be 0; the call fails)*/
  GRR_CNT_SLOTS = 16       /* length of a counters array (uint64)               */
};

const char *grr_last_error(void);
int grr_abi_version(void);

/* -- chain / task ------------------------------------------------------------------------- */
/* replaces RedundancyResolutionGraph.__init__ + readJsonSetup (graph.py:494-509,518-640) */
int grr_chain_create(const grr_chain_desc *desc, int device, grr_handle_t *out);
int grr_chain_destroy(grr_handle_t h);
/* replaces IKTask.__init__ orientation selection (graph.py:59-80); R_goal9 may be NULL */
int grr_task_set(grr_handle_t h, int mode, const double *R_goal9);
int grr_chain_num_joints(grr_handle_t h);
/* 'spin' joints (continuous rotation; Klamp't `joint spin`, flagged by the reference at graph.py:556-561): flags [host]
 * [n], non-zero = spin, NULL = none.  For such a joint robot.distance (solver.py:805, graph.py:946,955) takes the wrapped
 * angular difference, robot.interpolate (graph.py:962) the shorter arc normalised to [0, 2 pi), the IK solver treats it as
 * unbounded, draws its random starts from [0, 2 pi) and returns it normalised to [0, 2 pi).  Call before any compute call. */
int grr_chain_set_spin_joints(grr_handle_t h, const int32_t *flags);

/* -- K: forward kinematics (Klamp't robot.setConfig + link.getTransform; graph.py:747-756,887-893)
 * q [dev] Real [n][P]; out_p [dev] Real [P][3]; out_R [dev] Real [P][9] (may be NULL). */
int grr_fk_batch(grr_handle_t h, int dtype, int64_t P, const void *q, void *out_p, void *out_R, void *stream);

/* -- V2/K3/K4: seeded IK solves (RedundancyResolutionGraph.solve, graph.py:779-790 ->
 * ikTemplate.solve).  targets [dev] Real [P][dw]; q [dev] Real [n][P] in: seeds, out: result;
 * status [dev] u8 [P]; iters [dev] u8 [P] (nullable); resid [dev] Real [P] max|F| (nullable);
 * counters [dev] u64[GRR_CNT_SLOTS] (nullable). */
int grr_ik_solve_batch(grr_handle_t h, int dtype, int64_t P, const void *targets, void *q, uint8_t *status,
                       uint8_t *iters, void *resid, uint64_t *counters, void *stream);

/* -- S2: per-node random-start sampling (parallelQSamplingAtP, solver.py:875-885; the random
 * phase of sampleConfigurationSpace, solver.py:817-821).  For node i, sample s: seed drawn from
 * Philox4x32-10(key=seed, ctr=(node_uid[i], s, dof/4, 0)) uniformly in [qmin,qmax], then one IK
 * solve toward params[i].  q_raw [dev] Real [Nw][n][Nqp]; status/iters [dev] u8 [Nw][Nqp]
 * (iters nullable); counters nullable.  sample_offset shifts s (incremental growth). */
int grr_ik_sample(grr_handle_t h, int dtype, int64_t Nw, int32_t Nq, int32_t Nqp, const void *params,
                  const int32_t *node_uid, uint64_t seed, int32_t sample_offset, void *q_raw,
                  uint8_t *status, uint8_t *iters, uint64_t *counters, void *stream);

/* -- S1: the reference's default sampleConfigurationSpace (solver.py:757-835) is a sequential sweep: node i
 * makes numAdjacent = int(|adj|/deg * Nq) attempts seeded from configs of its lower-numbered neighbours that
 * already have configs (adj), then Nq - numAdjacent + numFailed random-start attempts, and dedups in that order.
 * Node i depends only on lower-numbered neighbours, so the sweep runs as wavefronts over that DAG with an
 * identical result.  The random-start attempts depend on no other node: the caller computes all Nq of them for
 * every node first (grr_ik_sample, same Philox stream as S2), then per level:
 *   grr_ik_sample_level: seeded attempt s < numAdjacent of node node_list[l] starts from config r1 % count[j] of
 *     lower neighbour j = adj[r0 % |adj|], (r0,r1) = Philox(key=seed, ctr=(uid, sample_offset+s, 0, 1))
 *     (solver.py:794-799); results to level-local q_raw [dev] Real [L][n][Nqr], status [dev] u8 [L][Nqr] (caller
 *     pre-fills status with 255), nadj [dev] i32 [L] = numAdjacent.  lower_ptr [dev] i32 [Nw+1] / lower_idx: CSR of
 *     each node's neighbours with a smaller id in increasing order; degree [dev] i32 [Nw]; q_kept/count: roadmap so far.
 *   grr_dedup_sweep: greedy 1e-2 filter over the node's seeded results, then its first Nq - nadj + numFailed
 *     precomputed random results (q_rand [dev] Real [Nw][n][Nqr], status_rand [dev] u8 [Nw][Nqr]), appended to
 *     q_kept/count in place; adds the attempts the sweep ran / converged to counters. */
/* The whole sweep as ONE launch (a dataflow over the same DAG): persistent CTAs claim nodes in the order `order` [dev] i32
 * [Nw] (nodes sorted by level, level(i) = 1 + max level of the lower-numbered neighbours), wait on the neighbours' flags in
 * ready [dev] i32 [Nw] (zeroed by the call), run the node's seeded attempts, filter seeded + precomputed random-start
 * results (q_rand / status_rand as for grr_dedup_sweep) in the sweep's order into q_kept / count and release the node's
 * flag.  Result identical to the level-by-level calls below, which it replaces when a node's working set (its
 * candidates' distance bit matrix) fits shared memory; returns -5 otherwise.
 * biased != 0 (solver.py:780-785, `biased=True` on a roadmap that already holds start_total configs in nodes_with_configs
 * nodes): node i makes Nq * configuration_count / (nodes_with_configs * count_i) random-start attempts (none if it has
 * no configs), configuration_count being the running total of the sweep -- so `order` must be 0..Nw-1, node i also waits
 * for node i - 1, whose total it reads from total_after [dev] i64 [Nw], and runs its random-start attempts itself
 * (q_rand / status_rand are not read, Nqr = Nq); configs beyond the capacity Nqp_kept are dropped. */
int grr_ik_sample_sweep(grr_handle_t h, int dtype, int64_t Nw, const int32_t *order, int32_t Nq, int32_t Nqr, const void *params,
                        const int32_t *node_uid, const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree,
                        void *q_kept, int32_t Nqp_kept, int32_t *count, uint64_t seed, int32_t sample_offset,
                        int32_t seed_from_adjacent, const void *q_rand, const uint8_t *status_rand, double thresh, int32_t *ready,
                        uint64_t *counters, void *stream, int32_t biased, int64_t start_total, int32_t nodes_with_configs,
                        int64_t *total_after);
int grr_ik_sample_level(grr_handle_t h, int dtype, int64_t L, const int32_t *node_list, int32_t Nq, int32_t Nqr,
                        const void *params, const int32_t *node_uid, const int32_t *lower_ptr, const int32_t *lower_idx,
                        const int32_t *degree, const void *q_kept, int32_t Nqp_kept, const int32_t *count, uint64_t seed,
                        int32_t sample_offset, int32_t seed_from_adjacent, void *q_raw, uint8_t *status, int32_t *nadj,
                        uint64_t *counters, void *stream);
int grr_dedup_sweep(grr_handle_t h, int dtype, int64_t L, const int32_t *node_list, int32_t Nq, int32_t Nqs, const void *q_seed,
                    const uint8_t *status_seed, const int32_t *nadj, int32_t seed_from_adjacent, int32_t Nqr, const void *q_rand,
                    const uint8_t *status_rand, double thresh, int32_t Nqp_kept, void *q_kept, int32_t *count,
                    uint64_t *counters, void *stream);

/* -- S3: greedy first-wins duplicate filter (solver.py:802-813,824-835,886-897).  Keeps sample s
 * iff status==OK and distance to every previously kept config of the node >= thresh.  Existing
 * configs (count_in[i] of them already at the front of q_kept's block; count_in nullable = 0)
 * take part in the test.  q_kept [dev] Real [Nw][n][Nqp_kept]; keep_idx [dev] i32 [Nw][Nq] (sample
 * index of each newly kept config, nullable); count [dev] i32 [Nw] out (total kept, capped at
 * Nqp_kept). */
int grr_dedup(grr_handle_t h, int dtype, int64_t Nw, int32_t Nq, int32_t Nqp, const void *q_raw,
              const uint8_t *status, double thresh, int32_t Nqp_kept, void *q_kept, const int32_t *count_in,
              int32_t *keep_idx, int32_t *count, void *stream);

/* -- E1/E3 + V1: all-pairs edge construction (connectConfigurationSpace ->
 * connectConfigurationSpaceOnEdge -> testAndConnect -> validEdgeLinear; solver.py:700-755,689-697,
 * 484-501; graph.py:938-1006).  For every workspace edge e=(iw,jw), iw<jw, and every pair
 * (a<count[iw], b<count[jw]) not both older than old_count (nullable), runs the validEdge
 * predicate and sets bit b of mask[e][a][b/32].  mask [dev] u32 [E][Nq][W], W=(Nq+31)/32, is
 * fully overwritten for the edges given (bits of both-old pairs are carried over when old_count is
 * given).  edge_stats [dev] u32 [E][2] = {pairs tested by this call, valid bits now stored for the
 * edge} (nullable); counters nullable.  Synchronises `stream` once (16-byte readback of the pair-test load)
 * before the main launch: dense roadmaps run one CTA per workspace edge with both node blocks staged in shared
 * memory, sparse ones (a few configs per node) one global queue of pairs served by persistent lanes; the handle
 * keeps E+1 words of device scratch for the queue offsets.  Nq <= 65535.  Env GRR_EDGES=dense|flat overrides the
 * choice (tests / profiling). */
int grr_edges_build(grr_handle_t h, int dtype, int64_t E, const int32_t *wedges, const void *params,
                    const void *q_kept, const int32_t *count, const int32_t *old_count, int32_t Nq,
                    int32_t Nqp, double epsilon, uint32_t *mask, uint32_t *edge_stats, uint64_t *counters,
                    void *stream);

/* The same with edge sets equal to those of IEEE double arithmetic (the reference's: Klamp't / KrisLibrary compute in
 * double) at single-precision speed.  First pass: grr_edges_build's single-precision kernels, which additionally check
 * every discrete decision of validEdgeLinear's interpolated solves (max|F| < tol, the line search's sufficient-decrease
 * / minimum-step / convergence-on-x tests, step clipping, the rounding of Ndivs, the leaf distance threshold;
 * graph.py:938-1006 + SURVEY A.3) against the single-precision error of the compared quantities and list the pairs in
 * which one of them is a near-tie.  Second pass: the listed pairs (a few per cent) are decided again in double
 * precision on the double-precision configs and their bits overwritten.  q_kept32 must be q_kept64 rounded to float
 * (same layout), params32 likewise.  counters additionally receive GRR_CNT_FLAGGED / FLIPPED / FLAG_DROPPED
 * (FLAG_DROPPED != 0: the handle's near-tie list overflowed -- more than a quarter of the pairs -- and the result is
 * NOT exact).  The margins are `scale` times the measured single/double differences (default 8;
 * grr_set_flag_scale). */
int grr_edges_build_exact(grr_handle_t h, int64_t E, const int32_t *wedges, const float *params32, const double *params64,
                          const float *q_kept32, const double *q_kept64, const int32_t *count, const int32_t *old_count,
                          int32_t Nq, int32_t Nqp, double epsilon, uint32_t *mask, uint32_t *edge_stats, uint64_t *counters,
                          void *stream);
int grr_set_flag_scale(grr_handle_t h, double scale);
/* How grr_edges_build_exact decides a chain's pairs: 1 = single-precision pass + double-precision recheck of the listed
 * near-ties (planar chains: every joint about +-e_y; measured 0 differing verdicts of 1.05e8 on BASELINE config 2, 2-4 % of
 * the pairs re-tested), 2 = the double-precision kernels throughout (every other chain: the interpolated solves of 6-/7-joint
 * arms pass near singular configurations where single- and double-precision iterates drift apart without any decision being
 * close, a quarter to a third of their pairs would have to be re-tested and the two-pass build costs what double precision
 * costs).  Env GRR_EXACT_MODE=recheck|double overrides (experiments). */
int grr_edges_exact_mode(grr_handle_t h);

/* Work estimate of grr_edges_build per workspace edge: sum over the pairs it will test of
 * Ndivs + 1 = ceil(robot.distance(qa,qb) / eps) + 1 (graph.py:944-946), i.e. interpolated IK solves +
 * leaf tests.  Used to cut the edge list into equal-work shards (multi-GPU) -- pair counts alone
 * mis-balance by ~15 % because Ndivs varies across the workspace.  cost [dev] u64 [E]. */
int grr_edges_cost(grr_handle_t h, int dtype, int64_t E, const int32_t *wedges, const void *q_kept, const int32_t *count,
                   const int32_t *old_count, int32_t Nqp, double epsilon, uint64_t *cost, void *stream);

/* Materialise (iq,jq) lists from the bitmask (the Gw.edge[iw][jw]['qlist'] sets, solver.py:458).
 * offsets [dev] i64 [E+1] = exclusive scan of edge_stats[:,1] (caller computes); out [dev] i32
 * [total][2] (iq,jq) in row-major pair order. */
int grr_edges_compact(int64_t E, int32_t Nq, const uint32_t *mask, const int64_t *offsets, int32_t *out,
                      void *stream);

/* -- V1 on explicit pairs (RedundancyResolutionGraph.validEdge, graph.py:1051-1053; backs
 * testAndConnect and the worker's 'test' / 'test-multi' tasks, solver.py:2386-2397).
 * qa,qb [dev] Real [n][P]; wa,wb [dev] Real [P][dw]; valid [dev] u8 [P];
 * info [dev] i32 [P][4] = {Ndivs, ik solves run, fail kind (0 none,1 IK,2 leaf distance), iters} nullable */
int grr_valid_edge_batch(grr_handle_t h, int dtype, int64_t P, const void *qa, const void *qb, const void *wa,
                         const void *wb, double epsilon, uint8_t *valid, int32_t *info, uint64_t *counters,
                         void *stream);

/* -- halo exchange support (multi-GPU): gather / scatter whole node blocks so that
 * the boundary nodes' (count, Q) travel as one contiguous NCCL message per peer.
 * idx [dev] i32 [K] node indices; packed [dev] Real [K][n][Nqp]. */
int grr_gather_node_blocks(int dtype, int64_t K, const int32_t *idx, int32_t n, int32_t Nqp, const void *q,
                           const int32_t *count, void *packed, int32_t *packed_count, void *stream);
int grr_scatter_node_blocks(int dtype, int64_t K, const int32_t *idx, int32_t n, int32_t Nqp, const void *packed,
                            const int32_t *packed_count, void *q, int32_t *count, void *stream);

/* The exchange itself for hosts that do not use PyTorch (replaces the pickled qlists the reference's
 * master/worker pool sends over OS pipes, grr/solver.py:2298-2337,2367-2400).  In the node-blocked layout a node range
 * [node_lo, node_hi) of q [dev] Real [Nw][n][Nqp] is one contiguous slab, so a halo is a list of slabs, each sent to or
 * received from one peer rank in place: ncclGroupStart; ncclSend / ncclRecv per slab (and of count[node_lo..node_hi)
 * [dev] i32 when count != NULL); ncclGroupEnd, all on `stream`.  nccl_comm is the caller's ncclComm_t; the library
 * resolves the NCCL entry points from the libnccl.so.2 already loaded in the process (env GRR_NCCL_LIB overrides the
 * name), so it carries no link-time dependency on NCCL.  Returns -4 on an NCCL error. */
int grr_halo_exchange(grr_handle_t h, void *nccl_comm, int dtype, int32_t Nqp, int32_t n_slabs, const int32_t *peer,
                      const int32_t *is_send, const int64_t *node_lo, const int64_t *node_hi, void *q, int32_t *count,
                      void *stream);

/* -- the roadmap's consumer on the bitmasks: the constraint satisfaction problem of
 * RedundancyResolver.makeCSP (grr/solver.py:984-1022) -- one variable per workspace node with configs, one table
 * constraint per workspace edge whose C-space edge set is non-empty; the table is the edge's Nq x Nq bitmask --
 * and CSP.randomDescentAssignment / numIncompatible / numConflicts (grr/utils/csp.py:934-1041, :133-150, :79-84).
 * G assignments ("guesses", solver.py:2048-2056) are handled at once: assignment [dev] i32 [G][Nw], -1 = no
 * variable / unassigned.  Adjacency in CSR form over ALL workspace edges: adj_ptr [dev] i32 [Nw+1],
 * adj_node [dev] i32 [2E], adj_edge [dev] i32 [2E] = e if the node is the first endpoint of workspace edge e
 * (wedges[e][0], the row index of the bitmask), ~e (bitwise not) if it is the second.
 * edge_active [dev] u8 [E] = 1 iff the edge's bitmask has a set bit (solver.py:1019).  Nq <= 65535. */

/* random.choice(domain) per variable (csp.py:950-955): Philox4x32-10, key = seed,
 * counter = (node, guess_offset + g, 0, 2); value = r0 % count[node]. */
int grr_csp_random_assignment(int32_t Nw, int32_t G, const int32_t *count, uint64_t seed, int32_t guess_offset,
                              int32_t *assignment, void *stream);

/* node_conflicts [dev] i32 [G][Nw] = numIncompatible(node, assignment[node]) (nullable);
 * total [dev] i32 [G] = numConflicts(assignment) (nullable; zeroed by the call). */
int grr_csp_conflicts(int32_t Nw, int32_t Nq, int32_t G, const int32_t *adj_ptr, const int32_t *adj_node,
                      const int32_t *adj_edge, const uint32_t *mask, const uint8_t *edge_active,
                      const int32_t *assignment, int32_t *node_conflicts, int32_t *total, void *stream);

/* One descent visit (csp.py:996-1030) of every node of node_list [dev] i32 [n_list] for every guess.  The nodes
 * of one call must be pairwise non-adjacent (one colour class of the workspace graph); then the call equals the
 * reference's sequential visits of those nodes in any order.  A node moves to the first value with the fewest
 * conflicts if that is strictly fewer than it has now.  The reference's `active` set: a node is visited iff
 * in_order [dev] u8 [G][Nw] (the caller's snapshot of `active` taken when the pass began, csp.py:993-995) is set;
 * active [dev] u8 [G][Nw] is cleared for every visited node and set for the assigned neighbours of every node
 * that moved (csp.py:1014-1020, :1027-1028).  changed [dev] i32 [G] += number of nodes that moved. */
int grr_csp_descent_sweep(int32_t n_list, const int32_t *node_list, int32_t Nw, int32_t Nq, int32_t G,
                          const int32_t *adj_ptr, const int32_t *adj_node, const int32_t *adj_edge, const uint32_t *mask,
                          const uint8_t *edge_active, const int32_t *count, int32_t *assignment,
                          const uint8_t *in_order, uint8_t *active, int32_t *changed, void *stream);

/* -- visibility-graph variant.  The reference interleaves the choice of the pairs to test (distance sorts,
 * union-find, early exits) with validEdge; validEdge is a pure function of the pair, so the predicate runs in batches on the
 * edge kernels and the calls below replay the reference's sequential choice logic on the verdict bits.
 *
 * Same-node pairs: a workspace "edge" (i, i) in the wedges of grr_edges_build / _exact denotes the pairs a > b of node i,
 * tested as validEdge(q_a, q_b, x_i, x_i) (constructSelfMotionManifold's all-pairs branch, grr/solver.py:645-655); bit
 * (a, b), a > b, of that edge's mask. */

/* Explicit pair list: list [dev] u64 = {number of entries, pad, then per entry {edge << 32 | a << 16 | b, 0}} with `edge`
 * an index into wedges; every listed pair is tested with validEdge (graph.py:938-1006) in `dtype` and bit (a, b) of
 * mask[edge] set or cleared (entries beyond `cap` are ignored and counted in GRR_CNT_FLAG_DROPPED).  Backs the k-nearest
 * branch of constructSelfMotionManifold (:635-643) and the CC-pair tests of testAndConnectQccs / parallelEdgeConnection. */
int grr_edges_test_listed(grr_handle_t h, int dtype, const int32_t *wedges, const void *params, const void *q_kept, int32_t Nq,
                          int32_t Nqp, double epsilon, uint32_t *mask, uint32_t *edge_stats, const uint64_t *list, uint64_t cap,
                          uint64_t *counters, void *stream);

/* constructSelfMotionManifold (grr/solver.py:623-669) for the Nn nodes of node_list [dev] i32 (NULL: nodes 0..Nn-1), one
 * warp per node, on self_mask [dev] u32 [Nw][Nq][W] (bit (i, j) = validEdge(q_i, q_j, x, x)).  knn = 0: all-pairs branch
 * (for i ascending the j < i of other components by increasing (distance, j)); knn > 0: the k nearest configs of the
 * node by (distance, j), i itself included (KDTree.knearest, grr/utils/kdtree.py:333-339).  Outputs, indexed by node:
 * label [Nw][Nq] = index of the config's component in `qccs` (components ordered by smallest member; -1 beyond count),
 * ncc [Nw], parents [Nw][Nq] = `qcc_parents` (depth-first predecessors in the merge forest, roots in index order,
 * neighbours in merge order, :661-668), members [Nw][Nq] = configs sorted by (component, id), cc_start [Nw][Nq+1] =
 * first position of each component in members, ntests [Nw] = validEdge calls the reference's loop makes (nullable).
 * grr_manifold_knn_list emits the pairs the knn branch may test ({node << 32 | i << 16 | j, 0}, list[0] += entries) for
 * grr_edges_test_listed on wedges[node] = (node, node).  Distances in IEEE double without contraction; spin_mask: bit j = chain
 * joint j is a 'spin' joint (wrapped angular difference, see grr_chain_set_spin_joints; these calls take no handle).  Nq <= 512. */
int grr_manifold_knn_list(int dtype, int64_t Nn, const int32_t *node_list, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                          const int32_t *count, int32_t knn, uint64_t *list, uint64_t cap, uint32_t spin_mask, void *stream);
int grr_manifold_components(int dtype, int64_t Nn, const int32_t *node_list, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                            const int32_t *count, const uint32_t *self_mask, int32_t knn, int32_t *label, int32_t *ncc,
                            int32_t *parents, int32_t *members, int32_t *cc_start, int32_t *ntests, uint32_t spin_mask, void *stream);

/* testAndConnectQccs (grr/solver.py:517-561) / parallelEdgeConnection (:949-982) on E workspace edges.
 * grr_qcc_candidates: for every CC pair (ci, cj) of every edge, cand [dev] i32 [(cc_off[e] + ci*ncc[jw] + cj) * 2T + slot]:
 * slots [0, T) = the pairs a << 16 | b with the T = max_tests smallest (robot.distance, a, b) (:534-546), slots [T, 2T) =
 * of the remaining pairs the T with the smallest (grr_pair_hash, a, b) -- a keyed hash stands in for random.sample (:551);
 * -1 = no such pair.  cc_off [dev] i64 [E+1] = exclusive scan of ncc[iw]*ncc[jw] (caller computes); node_uid as in
 * grr_ik_sample (nullable: node index).
 * grr_qcc_wave: appends to `list` (format of grr_edges_test_listed) slots [slot0, slot1) of every CC pair none of whose
 * earlier slots has its bit set in tested_mask [dev] u32 [E][Nq][W] (the mask grr_edges_test_listed writes).
 * grr_qcc_connect: the reference's loops on the verdict bits; mode 0 = testAndConnectQccs (a component of node i stops at
 * the first component of node j it connects to, :548-550), mode 1 = parallelEdgeConnection (every CC pair).  Sets bit
 * (a, b) of mask [dev] u32 [E][Nq][W] for every connection; conn [dev] i32 [cc_off[E]] = a << 16 | b, -1 (none) or -2 (not
 * reached); nconn [dev] i32 [E] = numConnected; ntests [dev] i32 [E] = validEdge calls of the reference's loops (nullable).
 * swap = 0: the outer loop runs over the components of wedges[e][0] (connectConfigurationSpace calls testAndConnectQccs(iw, jw)
 * with iw < jw, :725-726); swap = 1: over those of wedges[e][1] (sampleConfigurationSpace calls it with the node being
 * sampled first and its lower-numbered neighbour second, :842-848).  Candidates and conn then hold (outer config) << 16 |
 * (inner config); list entries, tested_mask and mask stay in the orientation of wedges. */
int grr_qcc_candidates(int dtype, int64_t E, const int32_t *wedges, int32_t n, int32_t Nq, int32_t Nqp, const void *q_kept,
                       const int32_t *node_uid, const int32_t *ncc, const int32_t *members, const int32_t *cc_start,
                       const int64_t *cc_off, int32_t max_tests, uint64_t seed, int32_t swap, int32_t *cand, uint32_t spin_mask, void *stream);
int grr_qcc_wave(int64_t E, const int32_t *wedges, const int32_t *ncc, const int64_t *cc_off, int32_t max_tests, const int32_t *cand,
                 int32_t Nq, const uint32_t *tested_mask, int32_t slot0, int32_t slot1, int32_t swap, uint64_t *list, uint64_t cap,
                 void *stream);
int grr_qcc_connect(int64_t E, const int32_t *wedges, const int32_t *ncc, const int64_t *cc_off, int32_t max_tests, const int32_t *cand,
                    int32_t Nq, const uint32_t *tested_mask, int32_t mode, int32_t swap, uint32_t *mask, int32_t *conn,
                    int32_t *nconn, int32_t *ntests, void *stream);
/* the hash of grr_qcc_candidates (host function; lets a caller or a test reproduce the sampled candidates) */
uint32_t grr_pair_hash(uint64_t seed, uint32_t iw, uint32_t jw, uint32_t a, uint32_t b);

/* -- measurement support: FP32 FMA peak microbenchmark (returns achieved TFLOP/s; synchronous). */
int grr_fp32_peak(int device, double *tflops_out);
/* double-precision FMA peak: denominator for the double-precision kernels (sampling; edge phase of general chains) */
int grr_fp64_peak(int device, double *tflops_out);
/* same with the packed sm_100 FFMA2 instruction (two FP32 FMAs per instruction per lane) */
int grr_fp32x2_peak(int device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* GRR_B200_H_ */
