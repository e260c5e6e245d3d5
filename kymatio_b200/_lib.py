"""ctypes binding of libscat_b200.so (C ABI declared in include/scat_b200.h).

There is NO fallback: if the shared library is missing or a call fails the
product path raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C kymatio_b200/csrc``.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCAT_B200_LIB", os.path.join(_HERE, "lib", "libscat_b200.so"))
CSRC_DIR = os.path.join(_HERE, "csrc")

_lib = None


class ScatB200Error(RuntimeError):
    pass


class PlanDesc2D(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int32), ("N", ctypes.c_int32), ("J", ctypes.c_int32), ("L", ctypes.c_int32),
                ("max_order", ctypes.c_int32), ("pre_pad", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol the
# header declares is exported.
_c = ctypes
SIGNATURES = {
    "scat_version": (_c.c_int, []),
    "scat_last_error": (_c.c_char_p, []),
    "scat_launch_count": (_c.c_uint64, []),
    "scat_timing_enable": (None, [_c.c_int]),
    "scat_timing_report": (_c.c_size_t, [_c.c_char_p, _c.c_size_t]),
    "scat_phase_prof_read": (_c.c_int, [_c.POINTER(_c.c_uint64), _c.c_int, _c.c_int]),
    "scat_plan2d_create": (_c.c_int, [_c.POINTER(PlanDesc2D), _c.POINTER(_c.c_void_p)]),
    "scat_plan2d_destroy": (None, [_c.c_void_p]),
    "scat_plan2d_info": (_c.c_int, [_c.c_void_p] + [_c.POINTER(_c.c_int32)] * 5),
    "scat_plan2d_const_bytes": (_c.c_size_t, [_c.c_void_p]),
    "scat_plan2d_bind": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_int32,
                                    _c.POINTER(_c.c_void_p), _c.c_int32, _c.c_void_p]),
    "scat_plan2d_workspace_bytes": (_c.c_size_t, [_c.c_void_p, _c.c_int64]),
    "scat_plan2d_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t,
                                       _c.c_int64, _c.c_void_p]),
    "scat_plan2d_forward_peers": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_int32,
                                             _c.c_void_p, _c.c_size_t, _c.c_int64, _c.c_void_p]),
    "scat_plan2d_forward_save": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_void_p,
                                            _c.c_size_t, _c.c_int64, _c.c_void_p]),
    "scat_plan2d_order1_mode": (_c.c_int32, [_c.c_void_p, _c.c_int32]),
    "scat_plan2d_order1_workspace_bytes": (_c.c_size_t, [_c.c_void_p, _c.c_int32, _c.c_int64]),
    "scat_plan2d_order1_forward": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64,
                                              _c.c_void_p]),
    "scat_plan2d_order1_backward": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_void_p, _c.c_size_t, _c.c_int64, _c.c_void_p]),
    "scat_plan2d_order2_channels": (_c.c_int32, [_c.c_void_p, _c.c_int32]),
    "scat_plan2d_order2_forward": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p]),
    "scat_plan2d_order2_backward": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                               _c.c_int64, _c.c_void_p]),
    "scat_fft2d_const_bytes": (_c.c_size_t, [_c.c_int32, _c.c_int32, _c.c_int32]),
    "scat_fft2d_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_fft2d_exec": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32,
                                   _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_pad2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64] + [_c.c_int32] * 7 + [_c.c_void_p]),
    "scat_cdgmm": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int32,
                              _c.c_int32, _c.c_void_p]),
    "scat_subsample_fourier2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32,
                                            _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_modulus": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_complex_from_real": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_real_part": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_fft1d_const_bytes": (_c.c_size_t, [_c.c_int32, _c.c_int32]),
    "scat_fft1d_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_fft1d_exec": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32,
                                   _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_pad1d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32,
                              _c.c_void_p]),
    "scat_subsample_fourier1d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32,
                                            _c.c_void_p]),
    "scat1d_split": (_c.c_int, [_c.c_int32, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32)]),
    "scat1d_tables_bytes": (_c.c_size_t, [_c.c_int32]),
    "scat1d_tables_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p]),
    "scat1d_fin_tables_bytes": (_c.c_size_t, [_c.c_int32]),
    "scat1d_fin_tables_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p]),
    "scat1d_col_prod": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                   _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_double, _c.c_void_p]),
    "scat1d_row_mod": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p, _c.c_int32, _c.c_double,
                                  _c.c_void_p]),
    "scat1d_col_fwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_double, _c.c_void_p]),
    "scat1d_rfft": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat1d_tile_max": (_c.c_int, []),
    "scat1d_tile": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                               _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_double,
                               _c.c_void_p]),
    "scat1d_finish": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int64,
                                 _c.c_int32, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_double, _c.c_void_p]),
    "scat1d_row_mod_t0": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p, _c.c_int32, _c.c_double,
                                     _c.c_void_p]),
    "scat1d_tile_t0": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                  _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_double, _c.c_void_p]),
    "scat1d_finish_global": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int64,
                                        _c.c_void_p, _c.c_int64, _c.c_void_p]),
    "scat1d_finseg_bytes": (_c.c_size_t, []),
    "scat_fft3d_const_bytes": (_c.c_size_t, [_c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32]),
    "scat_fft3d_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_fft3d_exec": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32,
                                   _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_modulus_rotation": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_compute_integrals": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_int32,
                                          _c.c_int32, _c.c_void_p]),
    "scat3d_supported": (_c.c_int, [_c.c_int32, _c.c_int32, _c.c_int32]),
    "scat3d_tables_bytes": (_c.c_size_t, [_c.c_int32, _c.c_int32, _c.c_int32]),
    "scat3d_tables_init": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat3d_rfft": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32,
                               _c.c_void_p]),
    "scat3d_col_prod": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32,
                                   _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat3d_plane": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p,
                                _c.c_int32, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat3d_col_fwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32,
                                  _c.c_void_p]),
    "scat_cdgmm_bcast": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int64,
                                    _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_subsample_fourier2d_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32,
                                                _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_subsample_fourier1d_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_int32, _c.c_int32,
                                                _c.c_void_p]),
    "scat_modulus_rotation_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                             _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_compute_integrals_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p,
                                              _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_modulus_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_void_p]),
    "scat_pad2d_bwd": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64] + [_c.c_int32] * 7 + [_c.c_void_p]),
    "scat_filters2d_spatial": (_c.c_int, [_c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p, _c.c_void_p]),
    "scat_filters2d_fold": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_filters3d_solid_harmonic": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_double, _c.c_int32,
                                                 _c.c_int32, _c.c_int32, _c.c_void_p]),
    "scat_filters3d_gaussian": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _c.c_int32,
                                           _c.c_void_p]),
}


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise ScatB200Error("building libscat_b200.so failed")
    return LIB_PATH


def load():
    """Load the shared library (once) and attach the C signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScatB200Error(
            f"{LIB_PATH} not found: the torch_b200 backend has no CPU or eager fallback. "
            "Build it with `make -C kymatio_b200/csrc` (needs nvcc).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().scat_last_error()
        raise ScatB200Error(msg.decode() if msg else f"libscat_b200 error {code}")


def launch_count():
    return int(load().scat_launch_count())


def timing_enable(on=True):
    load().scat_timing_enable(1 if on else 0)


def timing_report():
    """-> list of dicts {label, count, ms, bytes} aggregated per kernel label (synchronises)."""
    lib = load()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.scat_timing_report(buf, len(buf))
    rows = []
    for line in buf.value.decode().splitlines():
        label, cnt, ms, nbytes = line.split("\t")
        rows.append({"label": label, "count": int(cnt), "ms": float(ms), "bytes": float(nbytes)})
    return rows
