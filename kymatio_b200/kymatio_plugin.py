"""Plug ``torch_b200`` into an installed, UNMODIFIED kymatio.

    import kymatio_b200.kymatio_plugin as plugin
    plugin.install()
    from kymatio.torch import Scattering2D
    S = Scattering2D(J=3, shape=(256, 256), backend='torch_b200').cuda()     # or backend=plugin.backend2d

What ``install()`` does (no file of the reference is touched):
  1. registers the module ``kymatio.scattering2d.backend.torch_b200_backend`` (attribute ``backend``)
     in ``sys.modules`` so that the frontend's string route resolves it
     (kymatio/frontend/base_frontend.py:37-42);
  2. rebinds the name ``scattering2d`` that the torch frontend imported
     (kymatio/scattering2d/frontend/torch_frontend.py:4,98-99) to a dispatcher which runs the whole core
     as ONE call into libscat_b200.so when ``backend.name == 'torch_b200'`` and falls through to the
     reference core for every other backend.

``TorchB200Backend2D`` also implements every primitive of the backend protocol
(kymatio/scattering2d/core/scattering2d.py:3-9) as its own CUDA kernel, with the reference's checks
and error strings (kymatio/backend/torch_backend.py:102-219, kymatio/scattering2d/backend/
torch_backend.py:19-180), so the unchanged core can also be driven primitive by primitive
(``install(fused=False)``) and the reference's primitive tests can be pointed at it.  Like the
reference's own GPU-only backend (torch_skcuda) the primitives are CUDA-only and not differentiable;
gradients are provided by the fused path.
"""
import collections
import os
import ctypes
import importlib
import sys
import threading
import types
import warnings

import torch

from . import _lib
from .engine2d import Engine2D, _DTYPES

NAME = "torch_b200"


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


NOT_DIFFERENTIABLE = ("The torch_b200 backend does not propagate gradients through this primitive (the fused 2-D path and "
                      "the 1-D / 3-D backends are differentiable; the 2-D eager primitives are forward-only, like "
                      "kymatio's torch_skcuda backend). Call it under torch.no_grad() or use install(fused=True).")


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _cuda_check(x, differentiable=False):
    if not x.is_cuda:
        raise TypeError("The torch_b200 backend needs CUDA tensors. Use the torch backend for CPU tensors.")
    if not differentiable and _wants_grad(x):
        raise RuntimeError(NOT_DIFFERENTIABLE)        # never hand back a silently detached result


def _dtype_code(x):
    if x.dtype not in _DTYPES:
        raise TypeError("torch_b200 supports float32 and float64 tensors.")
    return _DTYPES[x.dtype]


class _FftTables:
    """Device-resident twiddle/permutation tables per (n0, n1, dtype, device) - torch tensors, so the
    library itself never allocates."""
    _cache = {}

    @classmethod
    def get(cls, n0, n1, ref):
        key = (n0, n1, ref.dtype, ref.device.index)
        buf = cls._cache.get(key)
        if buf is None:
            lib = _lib.load()
            code = _dtype_code(ref)
            nbytes = lib.scat_fft2d_const_bytes(n0, n1, code)
            if nbytes == 0:
                raise _lib.ScatB200Error(lib.scat_last_error().decode())
            with torch.cuda.device(ref.device):
                buf = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
                _lib.check(lib.scat_fft2d_init(buf.data_ptr(), n0, n1, code, _stream(ref)))
            cls._cache[key] = buf
        return buf


class Pad(object):
    """kymatio/scattering2d/backend/torch_backend.py:19-86 (reflect padding + trailing real axis)."""

    def __init__(self, pad_size, input_size):
        self.pad_size = list(pad_size)
        self.input_size = list(input_size)

    def __call__(self, x):
        _cuda_check(x)
        batch_shape, (M, N) = x.shape[:-2], x.shape[-2:]
        x = x.reshape((-1, M, N)).contiguous()
        t, b, l, r = self.pad_size
        out = torch.empty((x.shape[0], M + t + b, N + l + r), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_pad2d(x.data_ptr(), out.data_ptr(), x.shape[0], M, N, t, b, l, r,
                                              _dtype_code(x), _stream(x)))
        return out.reshape(batch_shape + out.shape[-2:] + (1,))


class _Primitives2D:
    """The compute primitives of the 2-D backend protocol as this library's kernels.  The generic checks
    (input_checks, contiguous_check, complex_check, ...), `_is_complex/_is_real` and the reshape helpers are NOT
    re-typed here: install() composes this mixin with the reference's own kymatio.backend.torch_backend.TorchBackend,
    so they are inherited unchanged."""
    name = NAME
    Pad = Pad

    # -- primitives -------------------------------------------------------------------------------------
    @classmethod
    def modulus(cls, x):
        cls.complex_contiguous_check(x)
        _cuda_check(x)
        out = torch.empty(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_modulus(x.data_ptr(), out.data_ptr(), out.numel(), _dtype_code(x), _stream(x)))
        return out

    @classmethod
    def _cdgmm_checks(cls, A, B):
        """The argument contract of the reference's cdgmm (kymatio/backend/torch_backend.py:181-219): same conditions,
        same exception types and messages, in the same order."""
        (cls.contiguous_check if cls._is_real(B) else cls.complex_contiguous_check)(B)
        cls.complex_contiguous_check(A)
        a_dev, b_dev = A.device, B.device
        rules = (
            (A.shape[-len(B.shape):-1] != B.shape[:-1], RuntimeError, "The filters are not compatible for multiplication."),
            (A.dtype is not B.dtype, TypeError, "Input and filter must be of the same dtype."),
            (b_dev.type == "cuda" and a_dev.type == "cuda" and a_dev.index != b_dev.index, TypeError,
             "Input and filter must be on the same GPU."),
            (b_dev.type == "cuda" and a_dev.type != "cuda", TypeError, "Input must be on GPU."),
            (b_dev.type == "cpu" and a_dev.type == "cuda", TypeError, "Input must be on CPU."),
        )
        for failed, exc, msg in rules:
            if failed:
                raise exc(msg)

    @classmethod
    def cdgmm(cls, A, B):
        cls._cdgmm_checks(A, B)
        _cuda_check(A)
        n = B.numel() // B.shape[-1]
        out = torch.empty_like(A)
        with torch.cuda.device(A.device):
            _lib.check(_lib.load().scat_cdgmm(A.data_ptr(), B.data_ptr(), out.data_ptr(), A.numel() // 2 // n, n,
                                              int(cls._is_complex(B)), _dtype_code(A), _stream(A)))
        return out

    @classmethod
    def subsample_fourier(cls, x, k):
        cls.contiguous_check(x)
        cls.complex_check(x)
        _cuda_check(x)
        n0, n1 = x.shape[-3], x.shape[-2]
        out = torch.empty(x.shape[:-3] + (n0 // k, n1 // k, 2), dtype=x.dtype, device=x.device)
        G = x.numel() // (n0 * n1 * 2)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_subsample_fourier2d(x.data_ptr(), out.data_ptr(), G, n0, n1, int(k),
                                                            _dtype_code(x), _stream(x)))
        return out

    @classmethod
    def _fft(cls, x, inverse):
        n0, n1 = x.shape[-3], x.shape[-2]
        tables = _FftTables.get(n0, n1, x)
        out = torch.empty_like(x)
        G = x.numel() // (n0 * n1 * 2)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_fft2d_exec(tables.data_ptr(), x.data_ptr(), out.data_ptr(), G, n0, n1,
                                                   int(inverse), _dtype_code(x), _stream(x)))
        return out

    @classmethod
    def rfft(cls, x):
        cls.contiguous_check(x)
        cls.real_check(x)
        _cuda_check(x)
        xc = torch.empty(x.shape[:-1] + (2,), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_complex_from_real(x.data_ptr(), xc.data_ptr(), x.numel(), _dtype_code(x),
                                                          _stream(x)))
        return cls._fft(xc, False)

    @classmethod
    def ifft(cls, x):
        cls.contiguous_check(x)
        cls.complex_check(x)
        _cuda_check(x)
        return cls._fft(x, True)

    @classmethod
    def irfft(cls, x):
        cls.contiguous_check(x)
        cls.complex_check(x)
        _cuda_check(x)
        y = cls._fft(x, True)
        out = torch.empty(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_real_part(y.data_ptr(), out.data_ptr(), out.numel(), _dtype_code(x), _stream(x)))
        return out



class _Layout2D:
    """The two pure-layout members of the 2-D protocol (kept out of _Primitives2D, which the 1-D / 3-D backends also
    build on: their unpad / stack come from the reference's TorchBackend1D / from _Primitives3D)."""

    @staticmethod
    def unpad(in_):
        # kymatio/scattering2d/backend/torch_backend.py:158-176
        in_ = in_[..., 1:-1, 1:-1, :]
        return in_.reshape(in_.shape[:-1])

    @staticmethod
    def stack(arrays):
        return torch.stack(arrays, -3)


# the backend classes exist once install() has composed the mixins with the reference's base classes
backend2d = backend1d = backend3d = backend = None


# ---------------------------------------------------------------------------------------------------
# 1-D backend (kymatio/scattering1d/backend/torch_backend.py) - eager primitives
# ---------------------------------------------------------------------------------------------------
class _Fft1dTables:
    _cache = {}

    @classmethod
    def get(cls, n, ref):
        key = (n, ref.dtype, ref.device.index)
        buf = cls._cache.get(key)
        if buf is None:
            lib = _lib.load()
            code = _dtype_code(ref)
            nbytes = lib.scat_fft1d_const_bytes(n, code)
            if nbytes == 0:
                raise _lib.ScatB200Error(lib.scat_last_error().decode())
            with torch.cuda.device(ref.device):
                buf = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
                _lib.check(lib.scat_fft1d_init(buf.data_ptr(), n, code, _stream(ref)))
            cls._cache[key] = buf
        return buf


class _DifferentiableEager:
    """Routes the generic primitives through kymatio_b200.ops_eager (kernel + hand-written adjoint) whenever a gradient
    is requested; otherwise the plain forward kernels of TorchB200Backend2D run."""

    @classmethod
    def _n_total(cls, x):
        return x.shape[-2]

    @classmethod
    def modulus(cls, x):
        if not _wants_grad(x):
            return super().modulus(x)
        from . import ops_eager
        cls.complex_contiguous_check(x)
        _cuda_check(x, True)
        return ops_eager.modulus(x)

    @classmethod
    def cdgmm(cls, A, B):
        if not _wants_grad(A):
            return super().cdgmm(A, B)
        from . import ops_eager
        cls._cdgmm_checks(A, B)
        _cuda_check(A, True)
        return ops_eager.Cdgmm.apply(A, B)

    @classmethod
    def _dfft(cls, x, inverse):
        from . import ops_eager
        _cuda_check(x, True)
        return ops_eager.FftN.apply(x, inverse, cls._fft, cls._n_total(x))

    @classmethod
    def rfft(cls, x):
        if not _wants_grad(x):
            return super().rfft(x)
        from . import ops_eager
        cls.contiguous_check(x)
        cls.real_check(x)
        return cls._dfft(ops_eager.FromReal.apply(x), False)

    @classmethod
    def ifft(cls, x):
        if not _wants_grad(x):
            return super().ifft(x)
        cls.contiguous_check(x)
        cls.complex_check(x)
        return cls._dfft(x, True)

    @classmethod
    def irfft(cls, x):
        if not _wants_grad(x):
            return super().irfft(x)
        from . import ops_eager
        cls.contiguous_check(x)
        cls.complex_check(x)
        return ops_eager.RealPart.apply(cls._dfft(x, True))


class _Primitives1D:
    """The 1-D specific compute primitives; average_global, unpad and the joint time-frequency reshapes (pad_frequency,
    swap_time_frequency, unpad_frequency, split_frequency_axis) are inherited from the reference's TorchBackend1D."""
    Pad = None

    @classmethod
    def subsample_fourier(cls, x, k):
        # kymatio/scattering1d/backend/torch_backend.py:19-48
        cls.complex_check(x)
        if _wants_grad(x):
            from . import ops_eager
            _cuda_check(x, True)
            return ops_eager.SubsampleFourier1d.apply(x, int(k))
        _cuda_check(x)
        x = x.contiguous()
        N = x.shape[-2]
        out = torch.empty(x.shape[:-2] + (N // k, 2), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_subsample_fourier1d(x.data_ptr(), out.data_ptr(), x.numel() // (2 * N), N,
                                                            int(k), _dtype_code(x), _stream(x)))
        return out

    @staticmethod
    def pad(x, pad_left, pad_right, mode="reflect"):
        # torch_backend.py:51-82 (only the reflect mode is used by the scattering frontends)
        if mode != "reflect":
            raise ValueError("torch_b200 implements reflect padding only.")
        if (pad_left >= x.shape[-1]) or (pad_right >= x.shape[-1]):
            raise ValueError("Indefinite padding size (larger than tensor).")
        if _wants_grad(x):
            from . import ops_eager
            _cuda_check(x, True)
            return ops_eager.Pad1d.apply(x, int(pad_left), int(pad_right))[..., None]
        _cuda_check(x)
        x = x.contiguous()
        N = x.shape[-1]
        out = torch.empty(x.shape[:-1] + (N + pad_left + pad_right,), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_pad1d(x.data_ptr(), out.data_ptr(), x.numel() // N, N, int(pad_left),
                                              int(pad_right), _dtype_code(x), _stream(x)))
        return out[..., None]

    @classmethod
    def _fft(cls, x, inverse):
        N = x.shape[-2]
        tables = _Fft1dTables.get(N, x)
        out, tmp = torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_fft1d_exec(tables.data_ptr(), x.data_ptr(), tmp.data_ptr(), out.data_ptr(),
                                                   x.numel() // (2 * N), N, int(inverse), _dtype_code(x), _stream(x)))
        return out

    @classmethod
    def cfft(cls, x):
        cls.contiguous_check(x)
        cls.complex_check(x)
        if _wants_grad(x):
            return cls._dfft(x, False)
        _cuda_check(x)
        return cls._fft(x, False)

    @staticmethod
    def stack(arrays, dim=2):
        return torch.stack(arrays, dim=dim)




# ---------------------------------------------------------------------------------------------------
# 3-D backend (kymatio/scattering3d/backend/torch_backend.py) - eager primitives
# ---------------------------------------------------------------------------------------------------
class _Fft3dTables:
    _cache = {}

    @classmethod
    def get(cls, shape, ref):
        key = (tuple(shape), ref.dtype, ref.device.index)
        buf = cls._cache.get(key)
        if buf is None:
            lib = _lib.load()
            code = _dtype_code(ref)
            nbytes = lib.scat_fft3d_const_bytes(shape[0], shape[1], shape[2], code)
            if nbytes == 0:
                raise _lib.ScatB200Error(lib.scat_last_error().decode())
            with torch.cuda.device(ref.device):
                buf = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
                _lib.check(lib.scat_fft3d_init(buf.data_ptr(), shape[0], shape[1], shape[2], code, _stream(ref)))
            cls._cache[key] = buf
        return buf


class _Primitives3D:
    Pad = None

    @staticmethod
    def stack(arrays, L):
        # torch_backend.py:73-76
        S = torch.stack(arrays, dim=1)
        return S.reshape((S.shape[0], S.shape[1] // (L + 1), (L + 1)) + S.shape[2:])

    @classmethod
    def _fft(cls, x, inverse):
        M, N, O = x.shape[-4], x.shape[-3], x.shape[-2]
        tables = _Fft3dTables.get((M, N, O), x)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_fft3d_exec(tables.data_ptr(), x.data_ptr(), out.data_ptr(),
                                                   x.numel() // (2 * M * N * O), M, N, O, int(inverse), _dtype_code(x),
                                                   _stream(x)))
        return out

    @classmethod
    def _n_total(cls, x):
        return x.shape[-4] * x.shape[-3] * x.shape[-2]

    @classmethod
    def cdgmm3d(cls, A, B):
        return cls.cdgmm(A, B)

    @staticmethod
    def modulus_rotation(x, module=None):
        # torch_backend.py:102-124
        if _wants_grad(x, module):
            from . import ops_eager
            _cuda_check(x, True)
            return ops_eager.ModulusRotation.apply(x, module)
        _cuda_check(x)
        x = x.contiguous()
        out = torch.empty(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
        prev = 0 if module is None else module.contiguous().data_ptr()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_modulus_rotation(x.data_ptr(), prev, out.data_ptr(), out.numel(),
                                                         _dtype_code(x), _stream(x)))
        return out

    @staticmethod
    def compute_integrals(input_array, integral_powers):
        # torch_backend.py:127-151 (the result takes torch's default dtype there, which is kept)
        if _wants_grad(input_array):
            from . import ops_eager
            _cuda_check(input_array, True)
            return ops_eager.ComputeIntegrals.apply(input_array, tuple(float(q) for q in integral_powers))
        _cuda_check(input_array)
        x = input_array.contiguous()
        B = x.shape[0]
        powers = torch.tensor([float(q) for q in integral_powers], dtype=torch.float32, device=x.device)
        acc = torch.empty((B, len(integral_powers)), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_compute_integrals(x.data_ptr(), acc.data_ptr(), B, x.numel() // max(B, 1),
                                                          powers.data_ptr(), len(integral_powers), _dtype_code(x),
                                                          _stream(x)))
        return acc.to(torch.get_default_dtype())



# ---------------------------------------------------------------------------------------------------
# fused dispatch
# ---------------------------------------------------------------------------------------------------
class _EngineCache:
    """Small thread-safe LRU of engines (class-level backends are shared by nn.DataParallel replica threads): a lookup
    holds the lock while the engine is built, so two threads never build the same plan twice; the least recently used
    entry is dropped when the cache is full - a live engine stays referenced by whoever is using it."""

    def __init__(self, capacity):
        self.capacity, self._d, self._lock = capacity, collections.OrderedDict(), threading.RLock()

    def get(self, key, build):
        with self._lock:
            if key in self._d:
                self._d.move_to_end(key)
                return self._d[key]
            value = build()
            self._d[key] = value
            while len(self._d) > self.capacity:
                self._d.popitem(last=False)
            return value

    def clear(self):
        with self._lock:
            self._d.clear()

    def __len__(self):
        return len(self._d)


_engines = _EngineCache(32)
_originals = {}
_tls = threading.local()            # .frontend3d: the HarmonicScattering3D module whose scattering() is running
_warned = set()


def _warn_fallback(what, why):
    """One warning per (transform, reason): a torch_b200 call that leaves the fused schedule for the per-primitive eager
    kernels is several times slower and must not do so silently."""
    if (what, why) not in _warned:
        _warned.add((what, why))
        warnings.warn("torch_b200: %s runs on the eager per-primitive kernels instead of the fused schedule (%s)" % (what, why),
                      RuntimeWarning, stacklevel=3)


def _fused_scattering2d(x, pad, unpad, backend_, J, L, phi, psi, max_order, out_type="array"):
    """Same signature and return value as kymatio/scattering2d/core/scattering2d.py:1-88."""
    from .autograd2d import scattering2d_apply
    if not x.is_cuda:
        raise TypeError("The torch_b200 backend needs CUDA tensors. Use the torch backend for CPU tensors.")
    pre_pad = not isinstance(pad, Pad)            # the frontend installs a lambda when pre_pad=True
    M, N = x.shape[-2:]
    phi_levels = [lvl for lvl in phi["levels"]]
    psi_levels = [lvl for p in psi for lvl in p["levels"]]
    if phi_levels[0].dtype is not x.dtype:
        raise TypeError("Input and filter must be of the same dtype.")
    if phi_levels[0].device != x.device:
        raise TypeError("Input and filter must be on the same GPU." if phi_levels[0].is_cuda else "Input must be on CPU.")
    key = (M, N, J, L, max_order, pre_pad, x.dtype, x.device.index)
    eng = _engines.get(key, lambda: Engine2D(M, N, J, L, max_order, pre_pad, x.dtype, x.device))
    eng.bind(phi_levels, psi_levels)
    S = scattering2d_apply(eng, x.reshape((-1, M, N)).contiguous())
    if out_type == "array":
        return S
    out, ch = [], 0

    def push(j, n, theta):
        nonlocal ch
        out.append({"coef": S[:, ch], "j": j, "n": n, "theta": theta})
        ch += 1

    push((), (), ())
    for n1, p1 in enumerate(psi):
        push((p1["j"],), (n1,), (p1["theta"],))
    if max_order >= 2:
        for n1, p1 in enumerate(psi):
            for n2, p2 in enumerate(psi):
                if p2["j"] > p1["j"]:
                    push((p1["j"], p2["j"]), (n1, n2), (p1["theta"], p2["theta"]))
    return out


_engines1d = _EngineCache(16)


def _fused_scattering1d(U_0, backend_, filters, log2_stride, average_local):
    """Generator with the signature and yield order of kymatio/scattering1d/core/scattering1d.py:2-107; the whole
    cascade runs as fused launches (engine1d.py) and every yielded 'coef' is a view of one (B, K, M) buffer."""
    from .engine1d import Engine1D
    phi, psi1 = filters[0], filters[1]
    psi2 = filters[2] if len(filters) > 2 else None
    Np = U_0.shape[-2]
    tensors = list(phi["levels"]) + [lv for p in psi1 for lv in p["levels"]]
    if psi2 is not None:
        tensors += [lv for p in psi2 for lv in p["levels"]]
    owner = getattr(_tls, "frontend1d", None)
    glob = (not average_local) and getattr(owner, "average", None) == "global"
    unav = (not average_local) and not glob                      # T=0: the modulus fields themselves
    key = (U_0.device.index, Np, int(log2_stride), psi2 is None, glob, unav) + tuple((t.data_ptr(), t._version) for t in tensors)
    eng = _engines1d.get(key, lambda: Engine1D(Np, log2_stride, phi, psi1, psi2, U_0.device, average_global=glob,
                                               unaveraged=unav))
    U0_hat = eng.rfft(U_0.reshape(-1, Np).contiguous())
    if unav:
        mods1, mods2 = eng.forward_unaveraged(U0_hat)
        where1, where2 = {}, {}
        for gi, g in enumerate(eng.groups):
            for i, n1 in enumerate(g["n1"]):
                where1[n1] = (mods1[gi], i)
                for ci, c in enumerate(g["children"]):
                    where2[(n1, c["n2"])] = (mods2[(gi, ci)], i)
        B = U0_hat.shape[0]
        for kind, n1, n2, ch in eng.order:
            if kind == "S0":
                yield {"coef": U_0, "j": (), "n": ()}
            elif kind == "S1":
                t, i = where1[n1]
                yield {"coef": t[:, i].reshape(B, 1, -1, 1), "j": (psi1[n1]["j"],), "n": (n1,)}
            else:
                t, i = where2[(n1, n2)]
                yield {"coef": t[:, i].reshape(B, 1, -1, 1), "j": (psi1[n1]["j"], psi2[n2]["j"]), "n": (n1, n2)}
        return
    S = eng.forward(U0_hat)
    B = S.shape[0]
    for kind, n1, n2, ch in eng.order:
        coef = S[:, ch].reshape(B, 1, eng.M, 1)
        if glob:
            coef = coef.contiguous()          # backend.average_global checks contiguity (a (B, 1, 1, 1) scalar per signal)
        if kind == "S0":
            yield {"coef": coef, "j": (), "n": ()}
        elif kind == "S1":
            yield {"coef": coef, "j": (psi1[n1]["j"],), "n": (n1,)}
        else:
            yield {"coef": coef, "j": (psi1[n1]["j"], psi2[n2]["j"]), "n": (n1, n2)}


def _fusable1d(U_0, backend_, filters, log2_stride, average_local):
    if getattr(backend_, "name", None) != NAME:
        return False
    if not average_local:
        # average_local=False is either average='global' (sum over time: scat1d_finish_global) or T=0 (the modulus field
        # of every path at its own resolution: scat1d_*_t0).  install() records the running frontend, which knows which.
        owner = getattr(_tls, "frontend1d", None)
        if owner is None or getattr(owner, "average", None) not in ("global", False):
            if torch.is_tensor(U_0) and U_0.is_cuda:
                _warn_fallback("Scattering1D", "scattering1d() was not called through the frontend: averaging mode unknown")
            return False
    if _wants_grad(U_0):
        return False                      # gradients: the unchanged core drives the differentiable eager primitives
    if not (torch.is_tensor(U_0) and U_0.is_cuda and U_0.dtype == torch.float32 and U_0.dim() == 4):
        return False
    if filters[0]["levels"][0].dtype != torch.float32 or not filters[0]["levels"][0].is_cuda:
        return False
    Np = U_0.shape[-2]
    jmax = max([p["j"] for f in filters[1:] for p in f] + [0])
    if not average_local:
        return Np & (Np - 1) == 0 and Np <= (1 << 18) and (Np >> jmax) >= 16
    M = Np >> int(log2_stride)
    return Np & (Np - 1) == 0 and Np <= (1 << 18) and 8 <= M <= 1024 and (Np >> max(
        [p["j"] for p in filters[1]] + [0])) >= 16


_engines3d = _EngineCache(16)


def _fused_scattering3d(x, filters, rotation_covariant, L, J, max_order, backend_, averaging):
    """Same signature and return value as kymatio/scattering3d/core/scattering3d.py:1-75, or None when the
    configuration is outside the fused kernels (the caller then runs the reference core on the eager primitives)."""
    from .engine3d import Engine3D, Unsupported
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5):
        return None
    if _wants_grad(x):
        return None                       # gradients: the unchanged core drives the differentiable eager primitives
    if any((not f.is_cuda) or f.dtype != torch.float32 or not f.is_contiguous() for f in filters[:L + 1]):
        return None
    # install() wraps the frontend's scattering() so that the running module is known here: `averaging` is
    # `lambda x: backend.compute_integrals(x, self.integral_powers)` (scattering3d/frontend/torch_frontend.py:70-71)
    owner = getattr(_tls, "frontend3d", None)
    if owner is None or getattr(owner, "method", "integral") != "integral":
        _warn_fallback("HarmonicScattering3D", "scattering3d() was not called through the frontend: integral_powers unknown")
        return None
    powers = [float(q) for q in owner.integral_powers]
    if not powers or len(powers) > 8 or max_order not in (1, 2):
        _warn_fallback("HarmonicScattering3D", "more than 8 integral powers" if len(powers) > 8 else "unsupported max_order")
        return None
    M, N, O = x.shape[1:4]

    def build():
        try:
            return Engine3D(M, N, O, x.device)
        except Unsupported:
            return False

    eng = _engines3d.get((x.device.index, M, N, O), build)
    if not eng:
        _warn_fallback("HarmonicScattering3D", "no fused kernel instance for a %d x %d x %d volume" % (M, N, O))
        return None
    U0_hat = eng.rfft(x.reshape(x.shape[:4]).contiguous())
    return eng.forward(U0_hat, filters, bool(rotation_covariant), int(L), int(J), int(max_order), powers)


def _compose_backends():
    """Build the backend classes: this library's compute primitives on top of the reference's OWN base classes, so that
    the checks, reshape helpers and the joint time-frequency reshapes are inherited, not restated."""
    global backend2d, backend1d, backend3d, backend
    if backend2d is not None:
        return
    from kymatio.backend.torch_backend import TorchBackend as RefBackend
    from kymatio.scattering1d.backend.torch_backend import TorchBackend1D as RefBackend1D
    from kymatio.scattering3d.backend.torch_backend import TorchBackend3D as RefBackend3D
    backend2d = type("TorchB200Backend2D", (_Layout2D, _Primitives2D, RefBackend), {"name": NAME, "Pad": Pad, "__doc__": _Primitives2D.__doc__})
    backend1d = type("TorchB200Backend1D", (_DifferentiableEager, _Primitives1D, _Primitives2D, RefBackend1D),
                     {"name": NAME, "Pad": None, "__doc__": _Primitives1D.__doc__})
    backend3d = type("TorchB200Backend3D", (_DifferentiableEager, _Primitives3D, _Primitives2D, RefBackend3D),
                     {"name": NAME, "Pad": None})
    backend = backend2d


def install(fused=True):
    """Register the backend module and (optionally) the fused core dispatcher. Idempotent."""
    import kymatio.scattering2d.frontend.torch_frontend as tf2d   # the unmodified reference
    _compose_backends()

    mod_name = "kymatio.scattering2d.backend.torch_b200_backend"
    if mod_name not in sys.modules:
        mod = types.ModuleType(mod_name)
        mod.backend = backend2d
        mod.__doc__ = "torch_b200 backend (provided by kymatio_b200)"
        sys.modules[mod_name] = mod
        setattr(importlib.import_module("kymatio.scattering2d.backend"), "torch_b200_backend", mod)

    mod1 = "kymatio.scattering1d.backend.torch_b200_backend"
    if mod1 not in sys.modules:
        m1 = types.ModuleType(mod1)
        m1.backend = backend1d
        m1.__doc__ = "torch_b200 1-D backend (provided by kymatio_b200)"
        sys.modules[mod1] = m1
        setattr(importlib.import_module("kymatio.scattering1d.backend"), "torch_b200_backend", m1)

    mod3 = "kymatio.scattering3d.backend.torch_b200_backend"
    if mod3 not in sys.modules:
        m3 = types.ModuleType(mod3)
        m3.backend = backend3d
        m3.__doc__ = "torch_b200 3-D backend (provided by kymatio_b200)"
        sys.modules[mod3] = m3
        setattr(importlib.import_module("kymatio.scattering3d.backend"), "torch_b200_backend", m3)

    if "scattering2d" not in _originals:
        _originals["scattering2d"] = tf2d.scattering2d
    reference_core = _originals["scattering2d"]
    _install_filter_hooks()

    import kymatio.scattering1d.frontend.base_frontend as bf1d
    if "scattering1d" not in _originals:
        _originals["scattering1d"] = bf1d.scattering1d
        _originals["frontend1d.scattering"] = bf1d.ScatteringBase1D.scattering
    reference_core1d = _originals["scattering1d"]
    frontend_scattering1d = _originals["frontend1d.scattering"]

    def scattering1d_with_owner(self, x):
        # the unmodified method, with the running module recorded for the dispatcher (average = 'global' or False)
        prev = getattr(_tls, "frontend1d", None)
        _tls.frontend1d = self
        try:
            return frontend_scattering1d(self, x)
        finally:
            _tls.frontend1d = prev
    scattering1d_with_owner.__wrapped__ = frontend_scattering1d
    bf1d.ScatteringBase1D.scattering = scattering1d_with_owner
    if fused:
        def dispatch1d(U_0, backend_, filters, log2_stride, average_local):
            if _fusable1d(U_0, backend_, filters, log2_stride, average_local):
                from .engine1d import Unsupported
                try:
                    gen = _fused_scattering1d(U_0, backend_, filters, log2_stride, average_local)
                    first = next(gen)
                except Unsupported as e:
                    _warn_fallback("Scattering1D", str(e) or "configuration outside the fused kernels")
                    return reference_core1d(U_0, backend_, filters, log2_stride, average_local)
                import itertools
                return itertools.chain([first], gen)
            return reference_core1d(U_0, backend_, filters, log2_stride, average_local)
        dispatch1d.__wrapped__ = reference_core1d
        bf1d.scattering1d = dispatch1d
    else:
        bf1d.scattering1d = reference_core1d

    import kymatio.scattering3d.frontend.torch_frontend as tf3d
    if "scattering3d" not in _originals:
        _originals["scattering3d"] = tf3d.scattering3d
        _originals["frontend3d.scattering"] = tf3d.HarmonicScatteringTorch3D.scattering
    reference_core3d = _originals["scattering3d"]
    frontend_scattering3d = _originals["frontend3d.scattering"]

    def scattering_with_owner(self, input_array):
        # the unmodified method, with the running module recorded for the fused dispatcher (integral_powers, method)
        prev = getattr(_tls, "frontend3d", None)
        _tls.frontend3d = self
        try:
            return frontend_scattering3d(self, input_array)
        finally:
            _tls.frontend3d = prev
    scattering_with_owner.__wrapped__ = frontend_scattering3d
    tf3d.HarmonicScatteringTorch3D.scattering = scattering_with_owner
    if fused:
        def dispatch3d(x, filters, rotation_covariant, L, J, max_order, backend, averaging):
            if getattr(backend, "name", None) == NAME:
                S = _fused_scattering3d(x, filters, rotation_covariant, L, J, max_order, backend, averaging)
                if S is not None:
                    return S
            return reference_core3d(x, filters=filters, rotation_covariant=rotation_covariant, L=L, J=J,
                                    max_order=max_order, backend=backend, averaging=averaging)
        dispatch3d.__wrapped__ = reference_core3d
        tf3d.scattering3d = dispatch3d
    else:
        tf3d.scattering3d = reference_core3d

    if fused:
        def dispatch(x, pad, unpad, backend_, J, L, phi, psi, max_order, out_type="array"):
            if getattr(backend_, "name", None) == NAME:
                return _fused_scattering2d(x, pad, unpad, backend_, J, L, phi, psi, max_order, out_type)
            return reference_core(x, pad, unpad, backend_, J, L, phi, psi, max_order, out_type)
        dispatch.__wrapped__ = reference_core
        tf2d.scattering2d = dispatch
    else:
        tf2d.scattering2d = reference_core
    return backend2d


def _gpu_filters(frontend):
    """Filter banks of torch_b200 frontends are synthesised on the GPU (filter_bank_gpu.py) unless SCAT_B200_GPU_FILTERS=0."""
    return (getattr(getattr(frontend, "backend", None), "name", None) == NAME and torch.cuda.is_available()
            and os.environ.get("SCAT_B200_GPU_FILTERS", "1") != "0")


def _install_filter_hooks():
    """Constructor path: ``create_filters`` of the 2-D / 3-D base frontends (kymatio/scattering2d/frontend/base_frontend.py:34-36,
    kymatio/scattering3d/frontend/base_frontend.py:25-30) is wrapped so that a frontend bound to this backend gets the same
    containers filled by the device-side synthesis; any other backend runs the reference's numpy code."""
    import kymatio.scattering2d.frontend.base_frontend as bf2d
    import kymatio.scattering3d.frontend.base_frontend as bf3d
    if "create_filters2d" not in _originals:
        _originals["create_filters2d"] = bf2d.ScatteringBase2D.create_filters
        _originals["create_filters3d"] = bf3d.ScatteringBase3D.create_filters
    ref2d, ref3d = _originals["create_filters2d"], _originals["create_filters3d"]

    def create_filters2d(self):
        if not _gpu_filters(self):
            return ref2d(self)
        from .filter_bank_gpu import filter_bank_2d_gpu
        bank = filter_bank_2d_gpu(self._M_padded, self._N_padded, self.J, self.L, as_numpy=True)
        self.phi, self.psi = bank["phi"], bank["psi"]
    create_filters2d.__wrapped__ = ref2d

    def create_filters3d(self):
        if not _gpu_filters(self):
            return ref3d(self)
        from .filter_bank_gpu import solid_harmonic_filter_bank_gpu, gaussian_filter_bank_gpu
        self.filters = solid_harmonic_filter_bank_gpu(self.M, self.N, self.O, self.J, self.L, self.sigma_0, as_numpy=True)
        self.gaussian_filters = gaussian_filter_bank_gpu(self.M, self.N, self.O, self.J + 1, self.sigma_0, as_numpy=True)
    create_filters3d.__wrapped__ = ref3d

    bf2d.ScatteringBase2D.create_filters = create_filters2d
    bf3d.ScatteringBase3D.create_filters = create_filters3d


def uninstall():
    if "create_filters2d" in _originals:
        import kymatio.scattering2d.frontend.base_frontend as bf2d
        import kymatio.scattering3d.frontend.base_frontend as bf3d
        bf2d.ScatteringBase2D.create_filters = _originals["create_filters2d"]
        bf3d.ScatteringBase3D.create_filters = _originals["create_filters3d"]
    if "scattering2d" in _originals:
        import kymatio.scattering2d.frontend.torch_frontend as tf2d
        tf2d.scattering2d = _originals["scattering2d"]
    if "scattering1d" in _originals:
        import kymatio.scattering1d.frontend.base_frontend as bf1d
        bf1d.scattering1d = _originals["scattering1d"]
        bf1d.ScatteringBase1D.scattering = _originals["frontend1d.scattering"]
    if "scattering3d" in _originals:
        import kymatio.scattering3d.frontend.torch_frontend as tf3d
        tf3d.scattering3d = _originals["scattering3d"]
        tf3d.HarmonicScatteringTorch3D.scattering = _originals["frontend3d.scattering"]
