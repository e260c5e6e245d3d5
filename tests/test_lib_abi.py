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


def test_host_only_entry_points():
    """Entry points that need no device: the finish-segment record layout shared with engine1d.py, the N = Na*Nb split
    and the support predicates of the fused 1-D / 3-D kernels."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libscat_b200.so not built")
    from kymatio_b200.engine1d import _FINSEG
    lib = _lib.load()
    assert lib.scat1d_finseg_bytes() == _FINSEG.itemsize == 64
    na, nb = ctypes.c_int32(), ctypes.c_int32()
    for n in range(4, 19):
        assert lib.scat1d_split(1 << n, ctypes.byref(na), ctypes.byref(nb)) == 0
        assert na.value * nb.value == 1 << n and nb.value >= 16 and nb.value >= na.value
    assert lib.scat1d_split(1 << 19, ctypes.byref(na), ctypes.byref(nb)) != 0 and lib.scat_last_error()
    assert lib.scat1d_split(1000, ctypes.byref(na), ctypes.byref(nb)) != 0
    assert lib.scat1d_tile_max() == 8192
    assert lib.scat3d_supported(128, 128, 128) == 1 and lib.scat3d_supported(64, 32, 32) == 1
    assert lib.scat3d_supported(12, 16, 20) == 0 and lib.scat3d_supported(32, 32, 16) == 0
    assert lib.scat1d_tables_bytes(1 << 17) > 0 and lib.scat1d_fin_tables_bytes(512) > 0
    assert lib.scat1d_fin_tables_bytes(4096) == 0 and lib.scat_last_error()


def test_engine3d_band_layout_matches_oracle_order():
    # n_j axis: first the J+1 first-order scales, then the (j1, j2 > j1) pairs in loop order (core/scattering3d.py:62-73)
    from kymatio_b200.engine3d import band_layout
    first, second, n = band_layout(L=2, J=2, max_order=2)
    assert first == {0: 0, 1: 1, 2: 2} and second == {(0, 1): 3, (0, 2): 4, (1, 2): 5} and n == 6
    first, second, n = band_layout(L=3, J=3, max_order=1)
    assert n == 4 and second == {}
