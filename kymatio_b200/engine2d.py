"""Host-side handle on the fused 2D engine (``scat_plan2d_*`` in include/scat_b200.h).

One ``Engine2D`` per (geometry, dtype, device); it owns the plan, the constant buffer
(twiddles, scramble tables, scrambled filter copies - a torch tensor, the library never
allocates) and re-binds the filters whenever the frontend's buffers change
(kymatio/scattering2d/frontend/torch_frontend.py:48-70 re-reads them on every call).
"""
import ctypes
import os

import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.float64: 1}


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class Engine2D:
    def __init__(self, M, N, J, L, max_order, pre_pad, dtype, device):
        if dtype not in _DTYPES:
            raise TypeError("torch_b200 supports float32 and float64 inputs, got %s" % dtype)
        self.lib = _lib.load()
        self.dtype, self.device = dtype, torch.device(device)
        desc = _lib.PlanDesc2D(int(M), int(N), int(J), int(L), int(max_order), int(bool(pre_pad)),
                               _DTYPES[dtype], 0)
        handle = ctypes.c_void_p()
        # the plan reads the SM count of, and opts its kernels into large shared memory on, the CURRENT device
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_create(ctypes.byref(desc), ctypes.byref(handle)))
        self._plan = handle
        vals = [ctypes.c_int32() for _ in range(5)]
        _lib.check(self.lib.scat_plan2d_info(self._plan, *[ctypes.byref(v) for v in vals]))
        self.Mp, self.Np, self.out_h, self.out_w, self.K = [v.value for v in vals]
        with torch.cuda.device(self.device):
            self._const = torch.empty(self.lib.scat_plan2d_const_bytes(self._plan), dtype=torch.uint8,
                                      device=self.device)
        self._bound_key = None
        self._filters = None
        self.geometry = dict(M=int(M), N=int(N), J=int(J), L=int(L), max_order=int(max_order), pre_pad=bool(pre_pad))

    def __del__(self):
        plan, self._plan = getattr(self, "_plan", None), None
        if plan:
            try:
                self.lib.scat_plan2d_destroy(plan)
            except Exception:
                pass

    # -- filters ---------------------------------------------------------------------
    def bind(self, phi_levels, psi_levels):
        """phi_levels: J tensors; psi_levels: flattened levels in registration order."""
        tensors = list(phi_levels) + list(psi_levels)
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self._bound_key:
            return
        for t in tensors:
            if t.dtype != self.dtype:
                raise TypeError("Input and filter must be of the same dtype.")
            if t.device != self.device:
                raise TypeError("Input and filter must be on the same GPU.")
            if not t.is_contiguous():
                raise RuntimeError("Tensors must be contiguous.")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_bind(
                self._plan, self._const.data_ptr(), _ptr_array(phi_levels), len(phi_levels),
                _ptr_array(psi_levels), len(psi_levels), ctypes.c_void_p(stream)))
        self._bound_key = key
        self._filters = (list(phi_levels), list(psi_levels))

    # -- forward ---------------------------------------------------------------------
    def workspace(self, batch):
        # a fresh tensor per call: torch's caching allocator makes this cheap and keeps the
        # buffer's lifetime stream-ordered (safe with several streams / DataParallel replicas)
        need = self.lib.scat_plan2d_workspace_bytes(self._plan, int(batch))
        with torch.cuda.device(self.device):
            return torch.empty(need, dtype=torch.uint8, device=self.device)

    def saved_u1_buffers(self, B):
        """Buffers for the first-order spectra a later backward pass needs (one per scale with children whose blocks are
        fused; None elsewhere), or None when they would not fit comfortably (the backward then recomputes them)."""
        g = self.geometry
        if g["max_order"] < 2 or os.environ.get("SCAT_B200_SAVE_U1", "1") == "0":
            return None
        shapes = [((B * g["L"], self.Mp >> j, self.Np >> j, 2) if (j < g["J"] - 1 and self.order1_mode(j)
                                                                   and self.order2_channels(j) > 0) else None)
                  for j in range(g["J"])]
        esz = 4 if self.dtype == torch.float32 else 8
        need = sum(esz * s[0] * s[1] * s[2] * s[3] for s in shapes if s is not None)
        if need == 0:
            return None
        # allocator counters, not cudaMemGetInfo: that driver call costs milliseconds and this runs on every training step
        total = torch.cuda.get_device_properties(self.device).total_memory
        if need > 0.4 * (total - torch.cuda.memory_allocated(self.device)):
            return None
        try:
            return [torch.empty(s, dtype=self.dtype, device=self.device) if s is not None else None for s in shapes]
        except torch.cuda.OutOfMemoryError:
            return None

    def forward_saving(self, x):
        """forward + the saved first-order spectra (scat_plan2d_forward_save) -> (out, list or None)."""
        B = x.shape[0]
        saved = self.saved_u1_buffers(B) if B > 0 else None
        if saved is None:
            return self.forward(x), None
        out = torch.empty((B, self.K, self.out_h, self.out_w), dtype=self.dtype, device=self.device)
        ws = self.workspace(B)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        arr = (ctypes.c_void_p * len(saved))(*[(t.data_ptr() if t is not None else None) for t in saved])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_forward_save(self._plan, x.data_ptr(), out.data_ptr(), arr, ws.data_ptr(),
                                                         ws.numel(), B, ctypes.c_void_p(stream)))
        return out, saved

    def forward(self, x, out=None, peer_ptrs=None, multicast_ptr=None):
        """x: (B, M, N) contiguous on self.device -> (B, K, out_h, out_w).

        out: optional preallocated contiguous result tensor (e.g. this rank's block of a symmetric-memory buffer);
        peer_ptrs: device addresses of the SAME block inside the output tensors of peer GPUs - every coefficient plane is
        then stored there as well by the kernels that produce it (scat_plan2d_forward_peers);
        multicast_ptr: the NVLS multicast address of the block instead - one multimem.st per value reaches every rank."""
        B = x.shape[0]
        if out is None:
            out = torch.empty((B, self.K, self.out_h, self.out_w), dtype=self.dtype, device=self.device)
        elif tuple(out.shape) != (B, self.K, self.out_h, self.out_w) or not out.is_contiguous() or out.dtype != self.dtype:
            raise ValueError("out must be a contiguous (B, K, out_h, out_w) tensor of the engine's dtype")
        if B == 0:
            return out
        ws = self.workspace(B)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            if multicast_ptr:
                arr = (ctypes.c_void_p * 1)(int(multicast_ptr))
                _lib.check(self.lib.scat_plan2d_forward_peers(
                    self._plan, x.data_ptr(), out.data_ptr(), arr, -1, ws.data_ptr(), ws.numel(), B, ctypes.c_void_p(stream)))
            elif peer_ptrs:
                arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
                _lib.check(self.lib.scat_plan2d_forward_peers(
                    self._plan, x.data_ptr(), out.data_ptr(), arr, len(peer_ptrs), ws.data_ptr(), ws.numel(), B,
                    ctypes.c_void_p(stream)))
            else:
                _lib.check(self.lib.scat_plan2d_forward(
                    self._plan, x.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), B,
                    ctypes.c_void_p(stream)))
        return out

    # -- second-order block (autograd building block) --------------------------------------
    def order2_channels(self, j1):
        return int(self.lib.scat_plan2d_order2_channels(self._plan, int(j1)))

    def order2_forward(self, j1, u1, batch):
        out = torch.empty((batch, self.order2_channels(j1), self.out_h, self.out_w), dtype=self.dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_order2_forward(self._plan, int(j1), u1.data_ptr(), out.data_ptr(), int(batch),
                                                           ctypes.c_void_p(stream)))
        return out

    def order2_backward(self, j1, u1, gout, batch):
        gu1 = torch.empty_like(u1)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_order2_backward(self._plan, int(j1), u1.data_ptr(), gout.data_ptr(),
                                                            gu1.data_ptr(), int(batch), ctypes.c_void_p(stream)))
        return gu1

    # -- first-order block (autograd building block) ----------------------------------------
    def order1_mode(self, j1):
        """0: not fused; 1: tile level (S1 and U1 from one kernel); 2: streaming level (three-kernel chain + low-pass)."""
        if os.environ.get("SCAT_B200_ORDER1_FUSED", "1") == "0":
            return 0
        return int(self.lib.scat_plan2d_order1_mode(self._plan, int(j1)))

    def order1_forward(self, j1, u0, batch, want_u1, want_s1=True):
        """u0: (batch, Mp, Np, 2) -> (S1 (batch, L, oh, ow) or None, U1 (batch*L, n0, n1, 2) or None)."""
        mode, L = self.order1_mode(j1), self.geometry["L"]
        n0, n1 = self.Mp >> j1, self.Np >> j1
        s1 = (torch.empty((batch, L, self.out_h, self.out_w), dtype=self.dtype, device=self.device)
              if (mode == 1 or want_s1) else None)
        u1 = torch.empty((batch * L, n0, n1, 2), dtype=self.dtype, device=self.device) if (want_u1 or mode == 2) else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.scat_plan2d_order1_forward(
                self._plan, int(j1), u0.data_ptr(), s1.data_ptr() if s1 is not None else None,
                u1.data_ptr() if u1 is not None else None, int(batch), ctypes.c_void_p(stream)))
        return s1, u1

    def order1_backward(self, j1, u0, gs1, gu1, batch):
        """-> gradient w.r.t. u0, (batch, Mp, Np, 2)."""
        gu0 = torch.zeros_like(u0)
        need = self.lib.scat_plan2d_order1_workspace_bytes(self._plan, int(j1), int(batch))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            ws = torch.empty(max(need, 16), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.scat_plan2d_order1_backward(
                self._plan, int(j1), u0.data_ptr(), gs1.data_ptr() if gs1 is not None else None,
                gu1.data_ptr() if gu1 is not None else None, gu0.data_ptr(), ws.data_ptr(), ws.numel(), int(batch),
                ctypes.c_void_p(stream)))
        return gu0

    # -- backward --------------------------------------------------------------------
    def backward(self, x, grad_out, saved_u1=None):
        """Gradient of ``forward`` w.r.t. x: the cascade is rebuilt on differentiable ops whose forward and
        adjoint kernels are this library's own (ops2d.py) and back-propagated; processed in batch chunks to
        bound the memory of the intermediates.  saved_u1 (forward_saving): the first-order spectra kept by the forward -
        the fused first-order blocks then do not run their forward kernels again."""
        from .ops2d import eager_scattering2d, _Recompute
        g = self.geometry
        phi, psi = self._filters
        pads = None
        if not g["pre_pad"]:
            t, l = (self.Mp - g["M"]) // 2, (self.Np - g["N"]) // 2
            pads = (t, self.Mp - g["M"] - t, l, self.Np - g["N"] - l)
        per_img = self.K_paths_bytes()
        B = x.shape[0]
        chunk = max(1, min(B, int((8 << 30) // max(1, per_img))))
        chunk = (B + (B + chunk - 1) // chunk - 1) // ((B + chunk - 1) // chunk)        # equal-sized chunks
        gx = torch.empty_like(x)
        L = g["L"]
        for b0 in range(0, B, chunk):
            with torch.enable_grad():
                xc = x[b0:b0 + chunk].detach().requires_grad_(True)
                _Recompute.active = True
                _Recompute.saved = None if saved_u1 is None else {
                    j: t[b0 * L:(b0 + chunk) * L] for j, t in enumerate(saved_u1) if t is not None}
                try:
                    y = eager_scattering2d(xc, g["J"], g["L"], g["max_order"], pads, phi, psi, eng=self)
                finally:
                    _Recompute.active = False
                    _Recompute.saved = None
                gx[b0:b0 + chunk] = torch.autograd.grad(y, xc, grad_out[b0:b0 + chunk])[0]
        return gx

    def K_paths_bytes(self):
        """Rough bytes of live intermediates per image in the rebuilt graph."""
        g, esz = self.geometry, (4 if self.dtype == torch.float32 else 8)
        J, L = g["J"], g["L"]
        fused = all(self.order1_mode(j) for j in range(J)) and (
            g["max_order"] < 2 or all(self.order2_channels(j) > 0 for j in range(J - 1)))
        if fused:
            # fused blocks keep one spectrum per first-order path (U1), its gradient, and the first-order backward workspace
            return sum(L * (self.Mp >> j) * (self.Np >> j) * 2 * esz for j in range(J)) * 5
        total = 0
        for j1 in range(J):
            n1 = (self.Mp >> j1) * (self.Np >> j1)
            total += L * n1 * 2 * esz * 6
            if g["max_order"] == 2:
                for j2 in range(j1 + 1, J):
                    total += L * L * ((self.Mp >> j1) * (self.Np >> j1) + 5 * (self.Mp >> j2) * (self.Np >> j2)) * 2 * esz
        return total
