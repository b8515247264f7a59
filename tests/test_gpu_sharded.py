"""Multi-GPU equality: with N >= 2 GPUs visible, N ranks (torchrun, NCCL) build sharded roadmaps that must equal the
single-GPU roadmap bit for bit -- counts, kept configs, edge masks, pair and valid counts -- for both exchange modes and
for the exact and the plain edge build (tools/check_sharded.py).  Skipped on a 1-GPU box; the partition / plan / slab
exchange logic itself is covered on CPU with gloo in tests/test_sharding.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_sharded_roadmap_equals_single_gpu(built):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (1 visible)")
    world = 2 if n < 4 else 4
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py")],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("sharded == single GPU") == 8
    assert out.stdout.count("sharded incremental == single GPU incremental") == 2
    assert out.stdout.count("sharded seeded == single GPU seeded") == 2


def test_halo_exchange_c_abi(built):
    """grr_halo_exchange (the collective entry point of the C ABI for hosts without PyTorch) with a communicator the
    host created itself through NCCL's C API: one rank (send-to-self inside the group) and, with >= 2 GPUs, a ring."""
    import torch
    env = dict(os.environ)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_halo_abi.py")], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "grr_halo_exchange ok: rank 0 of 1" in out.stdout, out.stderr[-3000:]
    n = torch.cuda.device_count()
    if n >= 2:
        world = 2 if n < 4 else 4
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                              "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_halo_abi.py")],
                             cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0 and out.stdout.count("grr_halo_exchange ok") == world, out.stderr[-3000:]
