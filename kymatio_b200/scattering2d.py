"""``Scattering2D`` - torch frontend over the fused sm_100a engine.

Drop-in for ``kymatio.torch.Scattering2D`` (kymatio/scattering2d/frontend/
torch_frontend.py:8-111 + base_frontend.py:8-36): same constructor arguments, the same
``tensor<n>`` buffers in the same order (state_dict compatible), the same checks and
error strings, the same output layout ``batch_shape + (K, M//2^J, N//2^J)`` for
``out_type='array'`` and the same list of dicts for ``out_type='list'``.  The per-path
loop of core/scattering2d.py:14-86 does not run here: ``forward`` is one call into
libscat_b200.so.  CUDA only - CPU tensors raise, as kymatio's own GPU-only backend does
(kymatio/scattering2d/backend/torch_skcuda_backend.py:68-69).
"""
import os

import torch
import torch.nn as nn

from .engine2d import Engine2D
from .filter_bank2d import filter_bank_2d, padded_size_2d

__all__ = ["Scattering2D"]

BACKEND_NAME = "torch_b200"


class Scattering2D(nn.Module):
    def __init__(self, J, shape, L=8, max_order=2, pre_pad=False, backend=BACKEND_NAME, out_type="array"):
        super().__init__()
        self.frontend_name = "torch"
        name = backend if isinstance(backend, str) else getattr(backend, "name", None)
        if name != BACKEND_NAME:
            # kymatio/frontend/base_frontend.py:37-50
            raise ImportError("The backend " + str(name) + " is not supported by this frontend; "
                              "kymatio_b200 only provides '" + BACKEND_NAME + "'.")
        self.backend = backend
        self.J, self.L, self.shape = J, L, tuple(shape)
        self.max_order, self.pre_pad, self.out_type = max_order, pre_pad, out_type
        self.build()
        self.create_filters()
        self.register_filters()
        self._engines = {}

    # -- construction (base_frontend.py:19-36) ------------------------------------------
    def build(self):
        M, N = self.shape
        if 2 ** self.J > M or 2 ** self.J > N:
            raise RuntimeError("The smallest dimension should be larger than 2^J.")
        self._M_padded, self._N_padded = padded_size_2d(M, N, self.J)

    def create_filters(self):
        # synthesised on the device when one is visible (filter_bank_gpu.py, milliseconds); the numpy bank (seconds at
        # 272 x 272) is the construction path of CPU-only hosts - both match the reference's filters to float32 rounding
        if torch.cuda.is_available() and os.environ.get("SCAT_B200_GPU_FILTERS", "1") != "0":
            from .filter_bank_gpu import filter_bank_2d_gpu
            filters = filter_bank_2d_gpu(self._M_padded, self._N_padded, self.J, self.L, as_numpy=True)
        else:
            filters = filter_bank_2d(self._M_padded, self._N_padded, self.J, self.L)
        self.phi, self.psi = filters["phi"], filters["psi"]

    def register_filters(self):
        # torch_frontend.py:23-43: phi levels first, then psi in list order; (m, n, 1) buffers
        n = 0
        for level in self.phi["levels"]:
            self.register_buffer("tensor" + str(n), torch.from_numpy(level).clone().unsqueeze(-1))   # own copy: the bank is cached
            n += 1
        for psi in self.psi:
            for level in psi["levels"]:
                self.register_buffer("tensor" + str(n), torch.from_numpy(level).clone().unsqueeze(-1))   # own copy: the bank is cached
                n += 1

    def load_filters(self):
        """Current buffers, split as (phi levels, flattened psi levels) - torch_frontend.py:48-70."""
        buffers = dict(self.named_buffers())
        n_phi = len(self.phi["levels"])
        n_psi = sum(len(p["levels"]) for p in self.psi)
        phi = [buffers["tensor" + str(n)] for n in range(n_phi)]
        psi = [buffers["tensor" + str(n_phi + n)] for n in range(n_psi)]
        return phi, psi

    # -- forward -----------------------------------------------------------------------
    def forward(self, x):
        # kymatio/frontend/torch_frontend.py:13-18, kymatio/backend/torch_backend.py:103-107
        if x is None:
            raise TypeError("The input should be not empty.")
        if torch.is_tensor(x) and not x.is_contiguous():
            raise RuntimeError("Tensors must be contiguous.")
        return self.scattering(x)

    def _engine(self, dtype, device):
        key = (dtype, device.index, self.max_order)
        eng = self._engines.get(key)
        if eng is None:
            M, N = (self._M_padded, self._N_padded) if self.pre_pad else self.shape
            eng = Engine2D(M, N, self.J, self.L, self.max_order, self.pre_pad, dtype, device)
            self._engines[key] = eng
        return eng

    def scattering(self, input):
        # checks: torch_frontend.py:73-89
        if not torch.is_tensor(input):
            raise TypeError("The input should be a PyTorch Tensor.")
        if len(input.shape) < 2:
            raise RuntimeError("Input tensor must have at least two dimensions.")
        if not input.is_contiguous():
            raise RuntimeError("Tensor must be contiguous.")
        if (input.shape[-1] != self.shape[-1] or input.shape[-2] != self.shape[-2]) and not self.pre_pad:
            raise RuntimeError("Tensor must be of spatial size (%i,%i)." % (self.shape[0], self.shape[1]))
        if (input.shape[-1] != self._N_padded or input.shape[-2] != self._M_padded) and self.pre_pad:
            raise RuntimeError("Padded tensor must be of spatial size (%i,%i)." % (self._M_padded, self._N_padded))
        if self.out_type not in ("array", "list"):
            raise RuntimeError("The out_type must be one of 'array' or 'list'.")
        if self.max_order not in (1, 2):
            raise RuntimeError("max_order must be 1 or 2.")
        if not input.is_cuda:
            raise TypeError("The torch_b200 backend runs on CUDA tensors only; use the torch backend "
                            "for CPU tensors.")

        phi, psi = self.load_filters()
        if phi[0].dtype is not input.dtype:
            raise TypeError("Input and filter must be of the same dtype.")
        if phi[0].device != input.device:
            if not phi[0].is_cuda:
                raise TypeError("Input must be on CPU.")
            raise TypeError("Input and filter must be on the same GPU.")

        batch_shape = input.shape[:-2]
        x = input.reshape((-1,) + input.shape[-2:])
        eng = self._engine(input.dtype, input.device)
        eng.bind(phi, psi)
        from .autograd2d import scattering2d_apply
        S = scattering2d_apply(eng, x)

        if self.out_type == "array":
            return S.reshape(batch_shape + S.shape[-3:])
        return self._as_list(S, batch_shape)

    def _as_list(self, S, batch_shape):
        # path order and meta of core/scattering2d.py:25-28,48-51,80-83
        out, ch = [], 0
        new_shape = batch_shape + S.shape[-2:]

        def push(j, n, theta):
            nonlocal ch
            out.append({"coef": S[:, ch].reshape(new_shape), "j": j, "n": n, "theta": theta})
            ch += 1

        push((), (), ())
        for n1, p1 in enumerate(self.psi):
            push((p1["j"],), (n1,), (p1["theta"],))
        if self.max_order == 2:
            for n1, p1 in enumerate(self.psi):
                for n2, p2 in enumerate(self.psi):
                    if p2["j"] > p1["j"]:
                        push((p1["j"], p2["j"]), (n1, n2), (p1["theta"], p2["theta"]))
        return out
