"""The C-ABI library loads and exports every symbol include/scat_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from kymatio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "scat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scat(?:1d|3d)?_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_all_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libscat_b200.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    lib.scat_version.restype = ctypes.c_int
    assert lib.scat_version() >= 100


def test_plan_geometry_without_gpu():
    """Plan creation is host-only up to the kernel-attribute calls; on a CPU-only box it must
    fail loudly (no silent fallback), on a GPU box it reports the reference's geometry."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libscat_b200.so not built")
    import torch
    lib = _lib.load()
    desc = _lib.PlanDesc2D(256, 256, 3, 8, 2, 0, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.scat_plan2d_create(ctypes.byref(desc), ctypes.byref(h))
    if not torch.cuda.is_available():
        assert rc != 0 and lib.scat_last_error()
        return
    assert rc == 0
    v = [ctypes.c_int32() for _ in range(5)]
    assert lib.scat_plan2d_info(h, *[ctypes.byref(x) for x in v]) == 0
    assert [x.value for x in v] == [272, 272, 32, 32, 217]
    lib.scat_plan2d_destroy(h)
