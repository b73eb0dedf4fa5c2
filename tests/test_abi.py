"""CPU: the C-ABI library builds, loads and exports every symbol include/racformer_ops.h declares.
No compute call is made here (no GPU); argument validation that happens before any CUDA call is exercised."""
import ctypes
import os
import re

from racformer_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "racformer_ops.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(racf_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_for_sm100a():
    path = build.build()
    assert os.path.exists(path)
    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 8
    lib = ctypes.CDLL(build.build())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in racformer_ops.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.racf_version() == 100
    assert _lib.status_string(0) == "ok"
    assert "num_point" in _lib.status_string(-4)


def test_argument_errors_are_codes_not_crashes():
    lib = _lib.load()
    hw = (ctypes.c_int * 2)(4, 4)
    feats = (ctypes.c_void_p * 1)(None)
    # null pointers
    assert lib.racf_msmv_forward(feats, hw, 1, None, None, 1, 64, 2, 1, 1, None, None) == -1
    one = ctypes.c_void_p(16)
    feats = (ctypes.c_void_p * 1)(16)
    # too many points (reference: msmv_sampling.cpp:159)
    assert lib.racf_msmv_forward(feats, hw, 1, one, one, 1, 64, 2, 1, 129, one, None) == -4
    # level count
    assert lib.racf_msmv_forward(feats, hw, 0, one, one, 1, 64, 2, 1, 1, one, None) == -2
    assert lib.racf_msmv_forward(feats, hw, 9, one, one, 1, 64, 2, 1, 1, one, None) == -2
    # bad dims
    assert lib.racf_msmv_forward(feats, hw, 1, one, one, 0, 64, 2, 1, 1, one, None) == -3
    # mmcv's im2col_step contract: batch % min(batch, step) == 0
    assert lib.racf_msda_forward(one, one, one, one, one, 6, 16, 1, 64, 1, 1, 1, 4, one, None) == -5
    assert lib.racf_msda_forward(None, one, one, one, one, 6, 16, 1, 64, 1, 1, 1, 4, one, None) == -1
