"""grr_halo_exchange through ctypes with a communicator made by the HOST (ncclCommInitRank via ctypes, no torch.distributed
NCCL group): every rank owns a contiguous node range, sends its boundary slab to the next rank and receives the previous
rank's.  Single process (world 1): send-to-self / recv-from-self inside one group.
usage: python tools/check_halo_abi.py            (1 GPU)
       torchrun --nproc-per-node N tools/check_halo_abi.py   (torch.distributed/gloo is used only to pass the NCCL id around)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import globalredundancyresolution_b200 as grr

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
torch.cuda.init(); torch.zeros(1, device="cuda")
nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), t.numpy().tobytes(), 128)
comm = C.c_void_p()
nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

pdef, res = grr.load_problem("planar_3R", "baseline_cfg1.json", device=local)
n, Nw, Nqp = res.chain.n, 16 * world, 8
for dt in (torch.float32, torch.float64):
    q = torch.zeros((Nw, n, Nqp), dtype=dt, device="cuda")
    count = torch.zeros(Nw, dtype=torch.int32, device="cuda")
    lo, hi = 16 * rank, 16 * (rank + 1)
    q[lo:hi] = (torch.arange(lo, hi, device="cuda", dtype=dt)[:, None, None] + 0.25 * torch.arange(n, device="cuda", dtype=dt)[None, :, None]
                + 0.001 * torch.arange(Nqp, device="cuda", dtype=dt)[None, None, :])
    count[lo:hi] = torch.arange(lo, hi, device="cuda", dtype=torch.int32) % 7
    if world == 1:
        q[8:16] = 0; count[8:16] = 0
        want_q, want_c = q[0:3].clone(), count[0:3].clone()
        res.chain.halo_exchange(comm, [(0, True, 0, 3), (0, False, 10, 13)], q, count)
        torch.cuda.synchronize()
        assert torch.equal(q[10:13], want_q) and torch.equal(count[10:13], want_c) and float(q[13:16].abs().sum()) == 0
    else:
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        plo = 16 * prv
        res.chain.halo_exchange(comm, [(nxt, True, hi - 4, hi), (prv, False, plo + 12, plo + 16)], q, count)
        torch.cuda.synchronize()
        want = (torch.arange(plo + 12, plo + 16, device="cuda", dtype=dt)[:, None, None] + 0.25 * torch.arange(n, device="cuda", dtype=dt)[None, :, None]
                + 0.001 * torch.arange(Nqp, device="cuda", dtype=dt)[None, None, :])
        assert torch.equal(q[plo + 12:plo + 16], want)
        assert torch.equal(count[plo + 12:plo + 16], torch.arange(plo + 12, plo + 16, device="cuda", dtype=torch.int32) % 7)
print("grr_halo_exchange ok: rank %d of %d" % (rank, world), flush=True)
nccl.ncclCommDestroy.argtypes = [C.c_void_p]
nccl.ncclCommDestroy(comm)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
