"""CPU-side checks of the C ABI boundary: the library builds for sm_100a, loads, and exports
exactly the entry points include/vdet_b200.h declares (no compute calls -- no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "vdet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vdet_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from vdetlib_b200 import build
    return build.build()


def test_header_matches_binding(lib_path):
    from vdetlib_b200 import _lib
    declared = _header_functions()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _header_functions():
        assert hasattr(lib, name), name
    lib.vdet_abi_version.restype = ctypes.c_int
    assert lib.vdet_abi_version() == 2
    lib.vdet_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.vdet_last_error(), bytes)


def test_library_is_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_binding_loads_and_reports_errors(lib_path):
    from vdetlib_b200 import _lib
    lib = _lib.load()
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = lib.vdet_temporal_maxpool(None, None, 0, 1, 8, 8, None, 4, -1e5, None)
    assert rc == _lib.ERR_INVALID
    assert "odd" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "temporal_maxpool")
    assert lib.vdet_nms_workspace_bytes(1000, 0) > 1000 * 40


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle, and CPU tensors are rejected."""
    import torch
    from vdetlib_b200 import ops
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vdetlib_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(dirpath, f)
                assert "kernel_double" not in src and "cuda_fake" not in src, os.path.join(dirpath, f)
    # nor do the tools: only tests/, smoke() and bench.py's CPU legs execute anything under oracle/
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            assert "oracle" not in open(os.path.join(ROOT, "tools", f)).read(), f
    with pytest.raises(TypeError):
        ops.iou_matrix(torch.zeros(2, 4), torch.zeros(2, 4))
    from vdetlib_b200.utils import cython_nms
    with pytest.raises(ValueError):
        cython_nms.nms(np.zeros((2, 5), np.float64), 0.3)      # dtype check precedes any device work


def test_host_copy_stream_threads_and_alignment():
    """vdet_host_copy_stream(_mt): plain host memory in and out (no CUDA call): every byte lands, nothing
    outside the destination range is touched, for any alignment and thread count."""
    import numpy as np
    from vdetlib_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 63, 64, 65, 4095, 4096, 100003, 1200007):
        src = rng.integers(0, 256, n + 64, dtype=np.uint8)
        for so, do in ((0, 0), (1, 3), (7, 16), (0, 61)):
            for th in (1, 2, 3, 8, 0):
                dst = np.full(n + 128, 0xAB, np.uint8)
                assert lib.vdet_host_copy_stream_mt(dst.ctypes.data + do, src.ctypes.data + so, n, th) == 0
                assert np.array_equal(dst[do:do + n], src[so:so + n]), (n, so, do, th)
                assert np.all(dst[:do] == 0xAB) and np.all(dst[do + n:] == 0xAB), (n, so, do, th)
    dst = np.zeros(64, np.uint8)
    assert lib.vdet_host_copy_stream(dst.ctypes.data, src.ctypes.data, 64) == 0 and np.array_equal(dst, src[:64])
    assert lib.vdet_host_copy_stream_mt(None, src.ctypes.data, 8, 2) == _lib.ERR_INVALID
