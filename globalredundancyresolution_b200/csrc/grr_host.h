// grr_host.h -- host-side glue shared by the C-ABI unit and the kernel instantiation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <limits>
#include <math.h>
#include "../../include/grr_b200.h"

namespace grr {

struct HostChain {
  int n, mode, nlinks_eps, max_iters, device;
  double tol, lambda;
  double Rp[GRR_MAX_JOINTS * 9], tp[GRR_MAX_JOINTS * 3], axis[GRR_MAX_JOINTS * 3];
  double qmin[GRR_MAX_JOINTS], qmax[GRR_MAX_JOINTS];
  double ee_local[3], R_tail[9], R_goal[9];
  int spin[GRR_MAX_JOINTS];      // 1: 'spin' joint (grr_chain_set_spin_joints)
  // near-tie margins of the exact edge build (grr_edges_build_exact), see flag_margins() in grr_fill.cuh
  double fl[10];
};

enum { FL_F = 0, FL_X = 1, FL_ALAM = 2, FL_REL = 3, FL_LEAF = 4, FL_ND = 5, FL_MAXIT = 6, FL_DD = 7, FL_CURV = 8 };

void set_error(const char *fmt, ...);
inline unsigned spin_mask_of(const HostChain &h) {
  unsigned m = 0;
  for (int j = 0; j < h.n; j++) if (h.spin[j]) m |= 1u << j;
  return m;
}

// planar-y chain: every joint has identity preceding rotation and axis +-e_y, goal-link frame = last joint frame, free
// orientation (the reference's planar_{2,3,5,10,20}R robots): the chains with the two-dimensional specialisation of the
// lane pass, and the ones whose exact edge build is the single-precision pass + recheck
inline bool is_planar_y_chain(const HostChain &h) {
  if (h.mode != GRR_FREE) return false;
  for (int j = 0; j < h.n; j++) if (h.spin[j]) return false;
  for (int k = 0; k < 9; k++) if (fabs(h.R_tail[k] - ((k % 4 == 0) ? 1.0 : 0.0)) > 1e-14) return false;
  for (int j = 0; j < h.n; j++) {
    const double *R = h.Rp + 9 * j, *a = h.axis + 3 * j;
    for (int k = 0; k < 9; k++) if (fabs(R[k] - ((k % 4 == 0) ? 1.0 : 0.0)) > 1e-14) return false;
    if (fabs(fabs(a[1]) - 1.0) > 1e-14 || fabs(a[0]) > 1e-14 || fabs(a[2]) > 1e-14) return false;
  }
  return true;
}

// grow-only device scratch of the handle (reallocation synchronises the device)
struct ScratchReq { void *ctx; unsigned long long *(*get)(void *ctx, size_t words); };
// double-precision inputs of the exact edge build's second pass
struct ExactArgs { const void *params64, *q64; };

// One table of launchers per (precision, residual rows).  `work` is a zero-initialised-by-the-launcher
// device counter from which persistent lanes pull their next instance.
struct Ops {
  int (*ik_sample)(const HostChain &, long long Nw, int Nq, int Nqp, const void *params, const int32_t *uid,
                   unsigned long long seed, int sample_offset, void *q_raw, uint8_t *status, uint8_t *iters,
                   unsigned long long *work, unsigned long long *counters, cudaStream_t st);
  int (*ik_level)(const HostChain &, long long Lnodes, const int32_t *node_list, int Nq, int Nqr, const void *params,
                  const int32_t *uid, const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree,
                  const void *Qk, int Nqk, const int32_t *count, unsigned long long seed, int sample_offset,
                  int seed_from_adjacent, void *q_raw, uint8_t *status, int32_t *nadj, unsigned long long *work,
                  unsigned long long *counters, cudaStream_t st);
  int (*ik_solve_batch)(const HostChain &, long long P, const void *targets, void *q, uint8_t *status, uint8_t *iters,
                        void *resid, unsigned long long *work, unsigned long long *counters, cudaStream_t st);
  int (*valid_edge_batch)(const HostChain &, long long P, const void *qa, const void *qb, const void *wa, const void *wb,
                          double epsilon, uint8_t *valid, int32_t *info, unsigned long long *work,
                          unsigned long long *counters, cudaStream_t st);
  int (*edges_build)(const HostChain &, long long E, const int32_t *wedges, const void *params, const void *q_kept,
                     const int32_t *count, const int32_t *old_count, int Nq, int Nqp, double epsilon, uint32_t *mask,
                     uint32_t *edge_stats, unsigned long long *work, const ScratchReq *scratch /* nullable */,
                     unsigned long long *counters, cudaStream_t st, const ExactArgs *exact /* nullable */);
  int (*ik_sweep)(const HostChain &, long long Nn, const int32_t *order, int Nq, int Nqr, const void *params, const int32_t *uid,
                  const int32_t *lower_ptr, const int32_t *lower_idx, const int32_t *degree, void *Qk, int Nqk, int32_t *count,
                  unsigned long long seed, int sample_offset, int seed_from_adjacent, const void *q_rand, const uint8_t *st_rand,
                  double thresh, int32_t *ready, unsigned long long *work, unsigned long long *counters, cudaStream_t st,
                  int biased, long long start_total, int nodes_with_configs, long long *total_after);
  int (*edges_listed)(const HostChain &, const int32_t *wedges, const void *params, const void *q_kept, int Nq, int Nqp,
                      double epsilon, uint32_t *mask, uint32_t *edge_stats, const unsigned long long *list,
                      unsigned long long cap, unsigned long long *work, unsigned long long *counters, cudaStream_t st);
};

const Ops *ops_f32_m3();
const Ops *ops_f32_m6();
const Ops *ops_f64_m3();
const Ops *ops_f64_m6();

}  // namespace grr
