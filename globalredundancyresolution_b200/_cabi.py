"""ctypes binding of libgrr_b200.so (include/grr_b200.h).

This is the ONLY compute path of the package: if the shared library is missing or a call fails the
error is raised -- there is no CPU or PyTorch fallback.  PyTorch is used for device memory and
streams only; every pointer handed to the library is `tensor.data_ptr()`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRR_LIB: load another build of the same library, e.g. the one with index checks (`make debug`)
LIB_PATH = os.environ.get("GRR_LIB") or os.path.join(_HERE, "libgrr_b200.so")

GRR_F32, GRR_F64 = 0, 1
GRR_FREE, GRR_FIXED, GRR_VARIABLE = 0, 1, 2
IK_OK, IK_STALL, IK_CONVX, IK_MAXIT = 0, 1, 2, 3
CNT_IK_SOLVES, CNT_IK_ITERS, CNT_IK_EVALS, CNT_PAIRS, CNT_VALID, CNT_IK_OK = 0, 1, 2, 3, 4, 5
CNT_FLAGGED, CNT_FLIPPED, CNT_FLAG_DROPPED, CNT_SLOTS = 6, 7, 8, 16
MAX_JOINTS = 32

# every symbol include/grr_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = [
    "grr_last_error", "grr_abi_version", "grr_chain_create", "grr_chain_destroy", "grr_task_set",
    "grr_chain_num_joints", "grr_chain_set_spin_joints", "grr_fk_batch", "grr_ik_solve_batch", "grr_ik_sample", "grr_ik_sample_level", "grr_dedup", "grr_dedup_sweep",
    "grr_edges_build", "grr_edges_build_exact", "grr_set_flag_scale", "grr_edges_exact_mode", "grr_edges_cost", "grr_edges_compact", "grr_valid_edge_batch", "grr_gather_node_blocks",
    "grr_scatter_node_blocks", "grr_halo_exchange", "grr_csp_random_assignment", "grr_csp_conflicts", "grr_csp_descent_sweep",
    "grr_fp32_peak", "grr_fp32x2_peak", "grr_fp64_peak",
    "grr_edges_test_listed", "grr_manifold_knn_list", "grr_manifold_components", "grr_qcc_candidates", "grr_qcc_wave",
    "grr_qcc_connect", "grr_pair_hash", "grr_ik_sample_sweep",
]


class GrrError(RuntimeError):
    pass


class ChainDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("mode", C.c_int32), ("nlinks_eps", C.c_int32), ("max_iters", C.c_int32),
        ("tol", C.c_double), ("lambda_", C.c_double),
        ("Rp", C.POINTER(C.c_double)), ("tp", C.POINTER(C.c_double)), ("axis", C.POINTER(C.c_double)),
        ("qmin", C.POINTER(C.c_double)), ("qmax", C.POINTER(C.c_double)),
        ("ee_local", C.c_double * 3), ("R_tail", C.c_double * 9), ("R_goal", C.c_double * 9),
    ]


_lib = None
LAUNCHES = 0      # kernels of this library launched since last reset (bench.py: gpu_launches)


def lib():
    """Load libgrr_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrrError(
            "libgrr_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C globalredundancyresolution_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    L.grr_last_error.restype = C.c_char_p
    L.grr_last_error.argtypes = []
    L.grr_abi_version.restype = C.c_int
    L.grr_chain_create.argtypes = [C.POINTER(ChainDesc), C.c_int, C.POINTER(vp)]
    L.grr_chain_destroy.argtypes = [vp]
    L.grr_task_set.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.grr_chain_num_joints.argtypes = [vp]
    L.grr_chain_set_spin_joints.argtypes = [vp, C.POINTER(C.c_int32)]
    L.grr_fk_batch.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp]
    L.grr_ik_solve_batch.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp, vp, vp, vp]
    L.grr_ik_sample.argtypes = [vp, C.c_int, i64, i32, i32, vp, vp, u64, i32, vp, vp, vp, vp, vp]
    L.grr_ik_sample_level.argtypes = [vp, C.c_int, i64, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, u64, i32, i32, vp, vp, vp,
                                      vp, vp]
    L.grr_dedup_sweep.argtypes = [vp, C.c_int, i64, vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, dbl, i32, vp, vp, vp, vp]
    L.grr_dedup.argtypes = [vp, C.c_int, i64, i32, i32, vp, vp, dbl, i32, vp, vp, vp, vp, vp]
    L.grr_edges_build.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp, vp, i32, i32, dbl, vp, vp, vp, vp]
    L.grr_edges_build_exact.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, i32, i32, dbl, vp, vp, vp, vp]
    L.grr_set_flag_scale.argtypes = [vp, dbl]
    L.grr_edges_exact_mode.argtypes = [vp]
    L.grr_ik_sample_sweep.argtypes = [vp, C.c_int, i64, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, u64, i32, i32, vp, vp, dbl, vp,
                                      vp, vp, i32, i64, i32, vp]
    L.grr_edges_test_listed.argtypes = [vp, C.c_int, vp, vp, vp, i32, i32, dbl, vp, vp, vp, u64, vp, vp]
    L.grr_manifold_knn_list.argtypes = [C.c_int, i64, vp, i32, i32, i32, vp, vp, i32, vp, u64, C.c_uint32, vp]
    L.grr_manifold_components.argtypes = [C.c_int, i64, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, C.c_uint32, vp]
    L.grr_qcc_candidates.argtypes = [C.c_int, i64, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, u64, i32, vp, C.c_uint32, vp]
    L.grr_qcc_wave.argtypes = [i64, vp, vp, vp, i32, vp, i32, vp, i32, i32, i32, vp, u64, vp]
    L.grr_qcc_connect.argtypes = [i64, vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.grr_pair_hash.argtypes = [u64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.grr_pair_hash.restype = C.c_uint32
    L.grr_edges_cost.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp, i32, dbl, vp, vp]
    L.grr_edges_compact.argtypes = [i64, i32, vp, vp, vp, vp]
    L.grr_valid_edge_batch.argtypes = [vp, C.c_int, i64, vp, vp, vp, vp, dbl, vp, vp, vp, vp]
    L.grr_gather_node_blocks.argtypes = [C.c_int, i64, vp, i32, i32, vp, vp, vp, vp, vp]
    L.grr_scatter_node_blocks.argtypes = [C.c_int, i64, vp, i32, i32, vp, vp, vp, vp, vp]
    L.grr_halo_exchange.argtypes = [vp, vp, C.c_int, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.grr_csp_random_assignment.argtypes = [i32, i32, vp, u64, i32, vp, vp]
    L.grr_csp_conflicts.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.grr_csp_descent_sweep.argtypes = [i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.grr_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.grr_fp32x2_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.grr_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name not in ("grr_last_error", "grr_pair_hash"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    global LAUNCHES
    if what not in ("grr_chain_create", "grr_task_set", "grr_fp32_peak", "grr_fp32x2_peak", "grr_fp64_peak", "grr_set_flag_scale", "grr_edges_exact_mode", "grr_pair_hash", "grr_chain_set_spin_joints"):
        # edges_build = k_edge_load + k_edges_build (+ k_edges_recheck in the exact build)
        LAUNCHES += {"grr_edges_build": 2, "grr_edges_build_exact": 3}.get(what, 1)
    if rc != 0:
        raise GrrError("%s failed (%d): %s" % (what, rc, lib().grr_last_error().decode()))


def _ptr(t):
    """Device (or None) pointer of a torch tensor."""
    if t is None:
        return None
    assert t.is_contiguous(), "C ABI buffers must be contiguous"
    return C.c_void_p(t.data_ptr())


def _stream(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(torch_dtype):
    import torch
    if torch_dtype == torch.float32:
        return GRR_F32
    if torch_dtype == torch.float64:
        return GRR_F64
    raise GrrError("unsupported dtype %r (float32 = product path, float64 = verification)" % (torch_dtype,))


def round_up4(nq):
    return (int(nq) + 3) // 4 * 4


class ChainHandle:
    """Opaque (device, chain) handle: grr_chain_create / grr_chain_destroy."""

    def __init__(self, chain_arrays, mode, R_goal=None, nlinks_eps=None, tol=1e-3, max_iters=50, lam=0.01, device=0, lazy=False):
        """chain_arrays: dict with Rp [n,9], tp [n,3], axis [n,3], qmin [n], qmax [n], ee_local [3], R_tail [9]
        (numpy float64; produced by robot.RobotChain.fold())."""
        self.n = int(chain_arrays["qmin"].shape[0])
        if not (1 <= self.n <= MAX_JOINTS):
            raise GrrError("grr_chain_create: n=%d outside 1..%d" % (self.n, MAX_JOINTS))
        self.mode = int(mode)
        self.device = int(device)
        self._keep = {k: np.ascontiguousarray(np.asarray(chain_arrays[k], dtype=np.float64))
                      for k in ("Rp", "tp", "axis", "qmin", "qmax", "ee_local", "R_tail")}
        sp = chain_arrays.get("spin") if hasattr(chain_arrays, "get") else None
        self.spin = np.zeros(self.n, dtype=np.int32) if sp is None else np.ascontiguousarray(np.asarray(sp, dtype=np.int32))
        self.nlinks_eps = int(nlinks_eps if nlinks_eps is not None else self.n)
        self.tol, self.max_iters, self.lam = float(tol), int(max_iters), float(lam)
        self._R_goal = np.eye(3).reshape(9) if R_goal is None else np.asarray(R_goal, dtype=np.float64).reshape(9).copy()
        self._flag_scale = None
        self._handle = None
        if not lazy:
            self._h            # create (and validate) now

    @property
    def _h(self):
        """The library handle, created on first use (describing a chain does not need the library: the CPU reference
        arm of bench.py builds problems without ever mapping libgrr_b200.so)."""
        if self._handle is None:
            L = lib()
            d = ChainDesc()
            d.n = self.n
            d.mode = self.mode
            d.nlinks_eps = self.nlinks_eps
            d.max_iters = self.max_iters
            d.tol = self.tol
            d.lambda_ = self.lam
            dp = C.POINTER(C.c_double)
            for k in ("Rp", "tp", "axis", "qmin", "qmax"):
                setattr(d, k, self._keep[k].ctypes.data_as(dp))
            d.ee_local[:] = list(self._keep["ee_local"].reshape(3))
            d.R_tail[:] = list(self._keep["R_tail"].reshape(9))
            d.R_goal[:] = list(self._R_goal)
            h = C.c_void_p()
            check(L.grr_chain_create(C.byref(d), self.device, C.byref(h)), "grr_chain_create")
            self._handle = h
            if self.spin.any():
                check(L.grr_chain_set_spin_joints(h, self.spin.ctypes.data_as(C.POINTER(C.c_int32))), "grr_chain_set_spin_joints")
            if self._flag_scale is not None:
                check(L.grr_set_flag_scale(h, float(self._flag_scale)), "grr_set_flag_scale")
        return self._handle

    @property
    def spin_mask(self):
        """bit j = chain joint j is a 'spin' joint (the visibility entry points take it as an argument)"""
        return int(sum(1 << j for j in range(self.n) if self.spin[j]))

    @property
    def dw(self):
        return 6 if self.mode == GRR_VARIABLE else 3

    def close(self):
        if getattr(self, "_handle", None):
            lib().grr_chain_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def task_set(self, mode, R_goal=None):
        rg = None
        if R_goal is not None:
            arr = (C.c_double * 9)(*[float(v) for v in np.asarray(R_goal).reshape(9)])
            rg = C.cast(arr, C.POINTER(C.c_double))
        self.mode = int(mode)
        if R_goal is not None:
            self._R_goal = np.asarray(R_goal, dtype=np.float64).reshape(9).copy()
        if self._handle is not None:
            check(lib().grr_task_set(self._handle, int(mode), rg), "grr_task_set")

    # ---- thin wrappers: torch tensors in, status checked --------------------------------------
    def fk_batch(self, q, out_p, out_R=None):
        P = q.shape[1]
        check(lib().grr_fk_batch(self._h, dtype_code(q.dtype), P, _ptr(q), _ptr(out_p), _ptr(out_R),
                                 _stream(q.device)), "grr_fk_batch")

    def ik_solve_batch(self, targets, q, status, iters=None, resid=None, counters=None):
        P = q.shape[1]
        check(lib().grr_ik_solve_batch(self._h, dtype_code(q.dtype), P, _ptr(targets), _ptr(q), _ptr(status),
                                       _ptr(iters), _ptr(resid), _ptr(counters), _stream(q.device)),
              "grr_ik_solve_batch")

    def ik_sample(self, params, node_uid, seed, q_raw, status, iters=None, counters=None, Nq=None, sample_offset=0):
        Nw, n, Nqp = q_raw.shape
        Nq = Nqp if Nq is None else int(Nq)
        check(lib().grr_ik_sample(self._h, dtype_code(q_raw.dtype), Nw, Nq, Nqp, _ptr(params), _ptr(node_uid),
                                  C.c_uint64(int(seed) & (2 ** 64 - 1)), int(sample_offset), _ptr(q_raw),
                                  _ptr(status), _ptr(iters), _ptr(counters), _stream(q_raw.device)),
              "grr_ik_sample")

    def ik_sample_level(self, node_list, Nq, params, node_uid, lower_ptr, lower_idx, degree, q_kept, count, seed,
                        sample_offset, seed_from_adjacent, q_raw, status, nadj, counters=None):
        L = node_list.shape[0]
        check(lib().grr_ik_sample_level(self._h, dtype_code(q_raw.dtype), L, _ptr(node_list), int(Nq), q_raw.shape[2],
                                        _ptr(params), _ptr(node_uid), _ptr(lower_ptr), _ptr(lower_idx), _ptr(degree),
                                        _ptr(q_kept), q_kept.shape[2], _ptr(count), C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                        int(sample_offset), 1 if seed_from_adjacent else 0, _ptr(q_raw), _ptr(status),
                                        _ptr(nadj), _ptr(counters), _stream(q_raw.device)), "grr_ik_sample_level")

    def ik_sample_sweep(self, order, Nq, params, node_uid, lower_ptr, lower_idx, degree, q_kept, count, seed, sample_offset,
                        seed_from_adjacent, q_rand, status_rand, thresh, ready, counters=None, biased=None):
        """The seeded sweep as one dataflow launch.  Returns False when a node's working set does not fit shared memory
        (the caller then runs the level-by-level launches).  biased = (start_total, nodes_with_configs, total_after int64
        [Nw] device tensor): the `biased=True` re-sampling of solver.py:780-785 (q_rand / status_rand are None)."""
        global LAUNCHES
        b = biased or (0, 0, None)
        rc = lib().grr_ik_sample_sweep(self._h, dtype_code(q_kept.dtype), order.shape[0], _ptr(order), int(Nq),
                                       q_rand.shape[2] if q_rand is not None else round_up4(Nq),
                                       _ptr(params), _ptr(node_uid), _ptr(lower_ptr), _ptr(lower_idx), _ptr(degree), _ptr(q_kept),
                                       q_kept.shape[2], _ptr(count), C.c_uint64(int(seed) & (2 ** 64 - 1)), int(sample_offset),
                                       1 if seed_from_adjacent else 0, _ptr(q_rand), _ptr(status_rand), float(thresh), _ptr(ready),
                                       _ptr(counters), _stream(q_kept.device), 1 if biased else 0, int(b[0]), int(b[1]), _ptr(b[2]))
        if rc == -5:
            return False
        check(rc, "grr_ik_sample_sweep")
        return True

    def dedup_sweep(self, node_list, Nq, q_seed, status_seed, nadj, seed_from_adjacent, q_rand, status_rand, thresh, q_kept,
                    count, counters=None):
        L = node_list.shape[0]
        check(lib().grr_dedup_sweep(self._h, dtype_code(q_seed.dtype), L, _ptr(node_list), int(Nq), q_seed.shape[2],
                                    _ptr(q_seed), _ptr(status_seed), _ptr(nadj), 1 if seed_from_adjacent else 0,
                                    q_rand.shape[2], _ptr(q_rand), _ptr(status_rand), float(thresh), q_kept.shape[2],
                                    _ptr(q_kept), _ptr(count), _ptr(counters), _stream(q_seed.device)), "grr_dedup_sweep")

    def dedup(self, q_raw, status, thresh, q_kept, count, Nq=None, count_in=None, keep_idx=None):
        Nw, n, Nqp = q_raw.shape
        Nq = Nqp if Nq is None else int(Nq)
        check(lib().grr_dedup(self._h, dtype_code(q_raw.dtype), Nw, Nq, Nqp, _ptr(q_raw), _ptr(status),
                              float(thresh), q_kept.shape[2], _ptr(q_kept), _ptr(count_in), _ptr(keep_idx),
                              _ptr(count), _stream(q_raw.device)), "grr_dedup")

    def edges_build(self, wedges, params, q_kept, count, Nq, epsilon, mask, edge_stats=None, counters=None,
                    old_count=None):
        E = wedges.shape[0]
        Nqp = q_kept.shape[2]
        check(lib().grr_edges_build(self._h, dtype_code(q_kept.dtype), E, _ptr(wedges), _ptr(params),
                                    _ptr(q_kept), _ptr(count), _ptr(old_count), int(Nq), Nqp, float(epsilon),
                                    _ptr(mask), _ptr(edge_stats), _ptr(counters), _stream(q_kept.device)),
              "grr_edges_build")

    def edges_build_exact(self, wedges, params32, params64, q32, q64, count, Nq, epsilon, mask, edge_stats=None, counters=None,
                          old_count=None):
        """Edge sets of double-precision arithmetic at single-precision speed (near-tie flags + recheck)."""
        import torch
        assert q32.dtype == torch.float32 and q64.dtype == torch.float64 and q32.shape == q64.shape
        assert params32.dtype == torch.float32 and params64.dtype == torch.float64
        E = wedges.shape[0]
        check(lib().grr_edges_build_exact(self._h, E, _ptr(wedges), _ptr(params32), _ptr(params64), _ptr(q32), _ptr(q64),
                                          _ptr(count), _ptr(old_count), int(Nq), q32.shape[2], float(epsilon), _ptr(mask),
                                          _ptr(edge_stats), _ptr(counters), _stream(q32.device)), "grr_edges_build_exact")

    def exact_mode(self):
        """'recheck' (single-precision pass + double-precision recheck) or 'double' (double-precision kernels): what
        grr_edges_build_exact does for this chain."""
        rc = lib().grr_edges_exact_mode(self._h)
        if rc < 0:
            check(rc, "grr_edges_exact_mode")
        return {1: "recheck", 2: "double"}[rc]

    def set_flag_scale(self, scale):
        self._flag_scale = float(scale)
        if self._handle is not None:
            check(lib().grr_set_flag_scale(self._handle, float(scale)), "grr_set_flag_scale")

    def halo_exchange(self, nccl_comm, slabs, q, count=None):
        """grr_halo_exchange: slabs = [(peer, is_send, node_lo, node_hi), ...]; nccl_comm = the caller's ncclComm_t as an
        integer / c_void_p (hosts without PyTorch; the torch path of sharding.exchange_halo does the same through
        torch.distributed)."""
        k = len(slabs)
        peer = (C.c_int32 * k)(*[int(s[0]) for s in slabs])
        snd = (C.c_int32 * k)(*[1 if s[1] else 0 for s in slabs])
        lo = (C.c_int64 * k)(*[int(s[2]) for s in slabs])
        hi = (C.c_int64 * k)(*[int(s[3]) for s in slabs])
        check(lib().grr_halo_exchange(self._h, C.c_void_p(int(nccl_comm) if not isinstance(nccl_comm, C.c_void_p) else nccl_comm.value),
                                      dtype_code(q.dtype), q.shape[2], k, peer, snd, lo, hi, _ptr(q), _ptr(count), _stream(q.device)),
              "grr_halo_exchange")

    def edges_cost(self, wedges, q_kept, count, epsilon, cost, old_count=None):
        E = wedges.shape[0]
        check(lib().grr_edges_cost(self._h, dtype_code(q_kept.dtype), E, _ptr(wedges), _ptr(q_kept), _ptr(count),
                                   _ptr(old_count), q_kept.shape[2], float(epsilon), _ptr(cost), _stream(q_kept.device)),
              "grr_edges_cost")

    def edges_test_listed(self, wedges, params, q_kept, Nq, epsilon, mask, plist, counters=None, edge_stats=None):
        """Test the listed pairs (plist: int64 device tensor {count, pad, entries...}) and set / clear their mask bits."""
        cap = (plist.shape[0] - 2) // 2
        check(lib().grr_edges_test_listed(self._h, dtype_code(q_kept.dtype), _ptr(wedges), _ptr(params), _ptr(q_kept), int(Nq),
                                          q_kept.shape[2], float(epsilon), _ptr(mask), _ptr(edge_stats), _ptr(plist), cap,
                                          _ptr(counters), _stream(q_kept.device)), "grr_edges_test_listed")

    def valid_edge_batch(self, qa, qb, wa, wb, epsilon, valid, info=None, counters=None):
        P = qa.shape[1]
        check(lib().grr_valid_edge_batch(self._h, dtype_code(qa.dtype), P, _ptr(qa), _ptr(qb), _ptr(wa), _ptr(wb),
                                         float(epsilon), _ptr(valid), _ptr(info), _ptr(counters),
                                         _stream(qa.device)), "grr_valid_edge_batch")


def edges_compact(mask, Nq, offsets, out):
    E = mask.shape[0]
    check(lib().grr_edges_compact(E, int(Nq), _ptr(mask), _ptr(offsets), _ptr(out), _stream(mask.device)),
          "grr_edges_compact")


def csp_random_assignment(count, seed, guess_offset, assignment):
    """assignment i32 [G][Nw] <- random.choice(domain) per variable (csp.py:950-955), -1 where count == 0."""
    G, Nw = assignment.shape
    check(lib().grr_csp_random_assignment(Nw, G, _ptr(count), int(seed) & (2 ** 64 - 1), int(guess_offset), _ptr(assignment),
                                          _stream(count.device)), "grr_csp_random_assignment")


def csp_conflicts(Nq, adj_ptr, adj_node, adj_edge, mask, edge_active, assignment, node_conflicts=None, total=None):
    """numIncompatible per node / numConflicts per guess (csp.py:133-150, :79-84) on the device bitmasks."""
    G, Nw = assignment.shape
    check(lib().grr_csp_conflicts(Nw, int(Nq), G, _ptr(adj_ptr), _ptr(adj_node), _ptr(adj_edge), _ptr(mask), _ptr(edge_active),
                                  _ptr(assignment), _ptr(node_conflicts), _ptr(total), _stream(mask.device)), "grr_csp_conflicts")


def csp_descent_sweep(node_list, Nq, adj_ptr, adj_node, adj_edge, mask, edge_active, count, assignment, in_order, active,
                      changed):
    """One descent visit (csp.py:996-1030) of the pairwise non-adjacent nodes of node_list, every guess."""
    G, Nw = assignment.shape
    check(lib().grr_csp_descent_sweep(node_list.shape[0], _ptr(node_list), Nw, int(Nq), G, _ptr(adj_ptr), _ptr(adj_node),
                                      _ptr(adj_edge), _ptr(mask), _ptr(edge_active), _ptr(count), _ptr(assignment),
                                      _ptr(in_order), _ptr(active), _ptr(changed), _stream(mask.device)),
          "grr_csp_descent_sweep")


def manifold_knn_list(q, count, Nq, knn, plist, node_list=None, spin_mask=0):
    """Pairs the k-nearest branch of constructSelfMotionManifold may test (solver.py:635-643) -> plist."""
    Nn = q.shape[0] if node_list is None else node_list.shape[0]
    cap = (plist.shape[0] - 2) // 2
    check(lib().grr_manifold_knn_list(dtype_code(q.dtype), Nn, _ptr(node_list), q.shape[1], int(Nq), q.shape[2], _ptr(q), _ptr(count),
                                      int(knn), _ptr(plist), cap, int(spin_mask), _stream(q.device)), "grr_manifold_knn_list")


def manifold_components(q, count, Nq, self_mask, knn, label, ncc, parents, members, cc_start, ntests=None, node_list=None, spin_mask=0):
    """constructSelfMotionManifold (solver.py:623-669) replayed on the self mask's verdict bits."""
    Nn = q.shape[0] if node_list is None else node_list.shape[0]
    check(lib().grr_manifold_components(dtype_code(q.dtype), Nn, _ptr(node_list), q.shape[1], int(Nq), q.shape[2], _ptr(q), _ptr(count),
                                        _ptr(self_mask), int(knn), _ptr(label), _ptr(ncc), _ptr(parents), _ptr(members),
                                        _ptr(cc_start), _ptr(ntests), int(spin_mask), _stream(q.device)), "grr_manifold_components")


def qcc_candidates(wedges, q, Nq, uid, ncc, members, cc_start, cc_off, max_tests, seed, swap, cand, spin_mask=0):
    check(lib().grr_qcc_candidates(dtype_code(q.dtype), wedges.shape[0], _ptr(wedges), q.shape[1], int(Nq), q.shape[2], _ptr(q),
                                   _ptr(uid), _ptr(ncc), _ptr(members), _ptr(cc_start), _ptr(cc_off), int(max_tests),
                                   int(seed) & (2 ** 64 - 1), int(swap), _ptr(cand), int(spin_mask), _stream(q.device)), "grr_qcc_candidates")


def qcc_wave(wedges, ncc, cc_off, max_tests, cand, Nq, tested_mask, slot0, slot1, swap, plist):
    cap = (plist.shape[0] - 2) // 2
    check(lib().grr_qcc_wave(wedges.shape[0], _ptr(wedges), _ptr(ncc), _ptr(cc_off), int(max_tests), _ptr(cand), int(Nq),
                             _ptr(tested_mask), int(slot0), int(slot1), int(swap), _ptr(plist), cap, _stream(wedges.device)), "grr_qcc_wave")


def qcc_connect(wedges, ncc, cc_off, max_tests, cand, Nq, tested_mask, mode, swap, mask, conn, nconn, ntests=None):
    check(lib().grr_qcc_connect(wedges.shape[0], _ptr(wedges), _ptr(ncc), _ptr(cc_off), int(max_tests), _ptr(cand), int(Nq),
                                _ptr(tested_mask), int(mode), int(swap), _ptr(mask), _ptr(conn), _ptr(nconn), _ptr(ntests),
                                _stream(wedges.device)), "grr_qcc_connect")


def pair_hash(seed, iw, jw, a, b):
    """The keyed hash that orders the sampled candidates of a CC pair (host function of the library)."""
    return int(lib().grr_pair_hash(int(seed) & (2 ** 64 - 1), int(iw), int(jw), int(a), int(b)))


def gather_node_blocks(idx, q, count, packed, packed_count):
    K = idx.shape[0]
    _, n, Nqp = q.shape
    check(lib().grr_gather_node_blocks(dtype_code(q.dtype), K, _ptr(idx), n, Nqp, _ptr(q), _ptr(count),
                                       _ptr(packed), _ptr(packed_count), _stream(q.device)),
          "grr_gather_node_blocks")


def scatter_node_blocks(idx, packed, packed_count, q, count):
    K = idx.shape[0]
    _, n, Nqp = q.shape
    check(lib().grr_scatter_node_blocks(dtype_code(q.dtype), K, _ptr(idx), n, Nqp, _ptr(packed),
                                        _ptr(packed_count), _ptr(q), _ptr(count), _stream(q.device)),
          "grr_scatter_node_blocks")


def fp32_peak(device=0):
    out = C.c_double(0.0)
    check(lib().grr_fp32_peak(int(device), C.byref(out)), "grr_fp32_peak")
    return out.value


def fp64_peak(device=0):
    out = C.c_double(0.0)
    check(lib().grr_fp64_peak(int(device), C.byref(out)), "grr_fp64_peak")
    return out.value


def fp32x2_peak(device=0):
    out = C.c_double(0.0)
    check(lib().grr_fp32x2_peak(int(device), C.byref(out)), "grr_fp32x2_peak")
    return out.value
