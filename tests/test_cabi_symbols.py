"""The C-ABI library loads on a CPU-only box and exports every symbol include/grr_b200.h declares.
No compute call is made here (no GPU)."""
import os
import re

import numpy as np
import pytest

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "grr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(grr_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    L = _cabi.lib()
    names = header_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), "libgrr_b200.so does not export %s" % n
    assert sorted(_cabi.EXPORTS) == names
    assert L.grr_abi_version() == 2


def test_chain_create_validates_arguments(built):
    r = grr.planar_robot(3)
    fold = r.fold(2, [1, 0, 0])
    h = _cabi.ChainHandle(fold, _cabi.GRR_FREE)
    assert _cabi.lib().grr_chain_num_joints(h._h) == 3
    h.task_set(_cabi.GRR_FIXED, np.eye(3))
    h.close()
    bad = dict(fold); bad["axis"] = fold["axis"] * 2.0
    with pytest.raises(_cabi.GrrError, match="not unit"):
        _cabi.ChainHandle(bad, _cabi.GRR_FREE)
    with pytest.raises(_cabi.GrrError, match="bad mode"):
        _cabi.ChainHandle(fold, 7)
    bad = dict(fold); bad["qmin"] = fold["qmax"] + 1
    with pytest.raises(_cabi.GrrError, match="qmin>qmax"):
        _cabi.ChainHandle(bad, _cabi.GRR_FREE)
    big = grr.planar_robot(40)
    with pytest.raises(_cabi.GrrError, match="outside"):
        _cabi.ChainHandle(big.fold(39, [1, 0, 0]), _cabi.GRR_FREE)


def test_compute_without_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pdef, res, rr = grr.make_program("planar_3R", "baseline_cfg1.json")
    with pytest.raises(_cabi.GrrError, match="no CUDA device"):
        rr.sampleConfigurationSpace(4)
    with pytest.raises(_cabi.GrrError, match="no CUDA device"):
        res.solve([0, 0, 0], [1.5, 0, 0.5])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "globalredundancyresolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "grr_oracle" not in txt, f


def test_halo_exchange_validates_arguments_without_a_gpu():
    """grr_halo_exchange is exported and refuses a null communicator before touching NCCL or the device."""
    import ctypes as C
    L = _cabi.lib()
    assert L.grr_halo_exchange(None, None, 0, 4, 1, None, None, None, None, None, None, None) == -1
    assert b"null handle" in L.grr_last_error()
