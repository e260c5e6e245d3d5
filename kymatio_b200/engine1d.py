"""Host engine of the fused 1-D scattering path (``scat1d_*`` in include/scat_b200.h).

It replaces the per-primitive loop of kymatio/scattering1d/core/scattering1d.py:40-107 (for
``average_local=True``) by a fixed schedule of fused launches.  Paths are grouped by first-order scale j1:

    S0                         finish(U0_hat, phi[0])
    group j1 (all n1 with that j1, k1 = min(j1, log2_stride), N1 = Np / 2^k1)
        col_prod  (U0_hat * psi1[n1], periodise 2^k1, first half of ifft)
        row_mod   (second half of ifft, modulus, first half of rfft)
        col_fwd   (second half of rfft)  -> U1_hat                         [only when the group has children]
        finish    (U1_hat * phi[k1], periodise, irfft)                     -> S1[n1]
        for every psi2[n2] with j2 > j1 (k2 = max(min(j2, log2_stride) - k1, 0), N2 = N1 / 2^k2):
            col_prod (U1_hat * psi2[n2].levels[k1], periodise 2^k2), row_mod(leaf), finish -> S2[n1, n2]

The engine owns only torch tensors (tables, pointer/support/channel arrays, workspaces); the arithmetic is in
libscat_b200.so.  There is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

SUPPORT_THRESHOLD = 1e-7       # filter bins below this fraction of the filter's maximum are not read
LOWPASS_THRESHOLD = 1e-9       # same for the low-pass


class Unsupported(Exception):
    """The configuration cannot run on the fused kernels (the caller then drives the eager primitives)."""


def circular_support(f, thr):
    """Smallest circular interval (start, len) holding every bin with |f| > thr * max|f|."""
    a = np.abs(np.asarray(f, dtype=np.float64)).ravel()
    n = a.size
    sig = np.flatnonzero(a > thr * a.max())
    if sig.size == 0:
        return 0, 0
    if sig.size == 1:
        return int(sig[0]), 1
    gaps = np.diff(np.concatenate([sig, [sig[0] + n]]))      # gap after each significant bin (circular)
    g = int(np.argmax(gaps))
    start = int(sig[(g + 1) % sig.size])
    length = n - int(gaps[g]) + 1
    return start, int(min(n, length))


def lowpass_bins(f, thr):
    """Number Fc of leading bins (0..Fc-1) of a symmetric low-pass needed so that every dropped bin f in [Fc, N-Fc]
    is below thr * max; multiple of 16, at most N/2 + 1."""
    a = np.abs(np.asarray(f, dtype=np.float64)).ravel()
    n = a.size
    sig = np.flatnonzero(a > thr * a.max())
    far = int(np.max(np.minimum(sig, n - sig))) if sig.size else 0
    return int(min(n // 2 + 1, (far + 1 + 15) // 16 * 16))


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _Tables:
    def __init__(self, device):
        self.device, self.path, self.fin = device, {}, {}

    def for_length(self, N):
        t = self.path.get(N)
        if t is None:
            lib = _lib.load()
            nbytes = lib.scat1d_tables_bytes(int(N))
            if nbytes == 0:
                raise Unsupported(lib.scat_last_error().decode())
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(lib.scat1d_tables_init(t.data_ptr(), int(N), _stream(self.device)))
            self.path[N] = t
        return t

    def for_lowpass(self, M):
        t = self.fin.get(M)
        if t is None:
            lib = _lib.load()
            nbytes = lib.scat1d_fin_tables_bytes(int(M))
            if nbytes == 0:
                raise Unsupported(lib.scat_last_error().decode())
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(lib.scat1d_fin_tables_init(t.data_ptr(), int(M), _stream(self.device)))
            self.fin[M] = t
        return t


def _split(N):
    lib = _lib.load()
    na, nb = ctypes.c_int32(), ctypes.c_int32()
    if lib.scat1d_split(int(N), ctypes.byref(na), ctypes.byref(nb)) != 0:
        raise Unsupported(lib.scat_last_error().decode())
    return na.value, nb.value


def schedule(Np, log2_stride, phi, psi1, psi2):
    """Pure host logic (testable without a GPU): the launch groups, channel map and yield order of the cascade.

    phi/psi1/psi2: the frontend's filter dictionaries (only 'j' and the number of levels are read here).
    Returns dict(M, K, groups, order) where order lists (kind, n1, n2, channel) in the reference's yield order
    (core/scattering1d.py:57-107) and channels follow the frontend's sort by (order, n)
    (frontend/base_frontend.py:150)."""
    ls = int(log2_stride)
    M = Np >> ls
    n1s = list(range(len(psi1)))
    pairs = []
    if psi2 is not None:
        for n1 in n1s:
            for n2 in range(len(psi2)):
                if psi2[n2]["j"] > psi1[n1]["j"]:
                    pairs.append((n1, n2))
    chan1 = {n1: 1 + n1 for n1 in n1s}
    chan2 = {p: 1 + len(n1s) + r for r, p in enumerate(sorted(pairs))}
    groups = []
    for j1 in sorted({psi1[n]["j"] for n in n1s}):
        members = [n for n in n1s if psi1[n]["j"] == j1]
        k1 = min(j1, ls)
        grp = dict(j1=j1, k1=k1, N1=Np >> k1, n1=members, chan=[chan1[n] for n in members], children=[])
        if psi2 is not None:
            for n2 in range(len(psi2)):
                j2 = psi2[n2]["j"]
                if j2 > j1:
                    s2 = min(j2, ls)
                    k2 = max(s2 - k1, 0)
                    grp["children"].append(dict(n2=n2, j2=j2, k2=k2, N2=(Np >> k1) >> k2, level=k1 + k2,
                                                chan=[chan2[(n, n2)] for n in members]))
        groups.append(grp)
    order = [("S0", None, None, 0)]
    for n1 in n1s:
        order.append(("S1", n1, None, chan1[n1]))
        if psi2 is not None:
            for n2 in range(len(psi2)):
                if (n1, n2) in chan2:
                    order.append(("S2", n1, n2, chan2[(n1, n2)]))
    return dict(M=M, K=1 + len(n1s) + len(pairs), groups=groups, order=order)


_FINSEG = np.dtype([("src_off", "<i8"), ("ss_g", "<i8"), ("ss_part", "<i8"), ("phi", "<u8"), ("chan", "<u8"),
                    ("which", "<i4"), ("nparts", "<i4"), ("N", "<i4"), ("Fc", "<i4"), ("NI", "<i4"), ("line0", "<i4")])


class Engine1D:
    """One engine per (device, filter-buffer identity, Np, log2_stride, averaging mode).

    average_global=True is the reference's ``average='global'`` (core/scattering1d.py with average_local=False followed by
    backend.average_global, frontend/base_frontend.py:137-138): every path is subsampled by its own 2^j and the output is
    the SUM over time of its modulus field = bin 0 of the spectrum the cascade already computes; the low-pass tail is
    replaced by scat1d_finish_global and the output is (B, K, 1)."""

    def __init__(self, Np, log2_stride, phi, psi1, psi2, device, average_global=False, unaveraged=False):
        self.lib = _lib.load()
        assert self.lib.scat1d_finseg_bytes() == _FINSEG.itemsize
        self.device = torch.device(device)
        self.Np, self.ls = int(Np), int(log2_stride)
        self.average_global = bool(average_global)
        # unaveraged = the reference's T=0: every path's modulus field at its own resolution 2^-j is an output
        # (core/scattering1d.py:75-76,104-105); same subsampling rule as 'global', no low-pass tail at all
        self.unaveraged = bool(unaveraged)
        if self.Np & (self.Np - 1):
            raise Unsupported("padded length must be a power of two")
        if self.average_global or self.unaveraged:
            # k1 = j1 and k2 = j2 - j1 (core/scattering1d.py:63,88-89 with average_local=False): a stride that never caps
            self.ls = max([p["j"] for p in psi1] + ([p["j"] for p in psi2] if psi2 is not None else []) + [0])
        sch = schedule(self.Np, self.ls, phi, psi1, psi2)
        self.M, self.K, self.order = sch["M"], sch["K"], sch["order"]
        if self.average_global or self.unaveraged:
            self.M = 1
        self.tables = _Tables(self.device)
        # transforms up to this length run as ONE launch with the whole path in shared memory (scat1d_tile)
        self.tile_max = self.lib.scat1d_tile_max() if os.environ.get("SCAT_B200_1D_TILE", "1") != "0" else 0
        with torch.cuda.device(self.device):
            self._keep = []                        # tensors the device arrays point into
            if self.average_global or self.unaveraged:
                # only bin 0 is needed: the leaves' pruned transform keeps its minimum of 16 bins, phi is never read
                self.fin_tab, self.phi_lv, self.Fc = None, None, [16] * (self.ls + 2)
            else:
                self.fin_tab = self.tables.for_lowpass(self.M)
                phi_lv = [lv.reshape(-1) for lv in phi["levels"]]
                self.phi_lv = phi_lv
                self.Fc = [lowpass_bins(lv.detach().cpu().numpy(), LOWPASS_THRESHOLD) for lv in phi_lv]
            self.chan0 = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.tables.for_length(self.Np)        # validates Np
            self.groups = []
            # per-signal workspace (complex elements): Y = largest group, U1 = every parent group, part = every leaf
            self.per_signal = dict(Y=0, U1=0, part=0)
            for g in sch["groups"]:
                N1, NI = g["N1"], len(g["n1"])
                if N1 < self.M:
                    raise Unsupported("first-order length below the output length")
                if N1 < 16:
                    raise Unsupported("first-order length below 16 samples")
                filt = [psi1[n]["levels"][0].reshape(-1) for n in g["n1"]]
                gd = dict(g)
                gd.update(NI=NI, tab=self.tables.for_length(N1), **self._filter_arrays(filt, self.Np))
                gd["chan_dev"] = torch.tensor(g["chan"], dtype=torch.int32, device=self.device)
                gd["tile"] = N1 <= self.tile_max
                gd["nparts"] = 1 if gd["tile"] else (_split(N1)[0] + 15) // 16
                if not gd["tile"]:
                    self.per_signal["Y"] = max(self.per_signal["Y"], NI * N1)
                if g["children"]:
                    gd["u1_off"] = self.per_signal["U1"]
                    self.per_signal["U1"] += NI * N1
                else:
                    gd["part_off"] = self.per_signal["part"]
                    self.per_signal["part"] += NI * gd["nparts"] * self.Fc[g["k1"]]
                kids = []
                for c in g["children"]:
                    cd = dict(c)
                    if c["N2"] < self.M:
                        raise Unsupported("second-order length below the output length")
                    f2 = psi2[c["n2"]]["levels"][g["k1"]].reshape(-1)
                    cd.update(tab=self.tables.for_length(c["N2"]), **self._filter_arrays([f2] * NI, N1, same=True))
                    cd["chan_dev"] = torch.tensor(c["chan"], dtype=torch.int32, device=self.device)
                    cd["tile"] = c["N2"] <= self.tile_max
                    cd["nparts"] = 1 if cd["tile"] else (_split(c["N2"])[0] + 15) // 16
                    if not cd["tile"]:
                        self.per_signal["Y"] = max(self.per_signal["Y"], NI * c["N2"])
                    cd["part_off"] = self.per_signal["part"]
                    self.per_signal["part"] += NI * cd["nparts"] * self.Fc[c["level"]]
                    kids.append(cd)
                gd["children"] = kids
                self.groups.append(gd)
            self._segs = {}
            torch.cuda.current_stream(self.device).synchronize()

    def _filter_arrays(self, filt, Npar, same=False):
        for f in filt:
            if f.dtype != torch.float32 or not f.is_cuda or f.numel() != Npar or not f.is_contiguous():
                raise Unsupported("filters must be contiguous float32 CUDA tensors on the parent grid")
        self._keep.extend(filt)
        host = [filt[0].detach().cpu().numpy()] * len(filt) if same else [f.detach().cpu().numpy() for f in filt]
        supp = [circular_support(h, SUPPORT_THRESHOLD) for h in host]
        return dict(
            filt_dev=torch.tensor([f.data_ptr() for f in filt], dtype=torch.int64, device=self.device),
            supp_dev=torch.tensor(supp, dtype=torch.int32, device=self.device).reshape(-1, 2).contiguous(),
            supp_len=[s[1] for s in supp])

    def _segments(self, nb):
        """Segment table of the finish launch for a chunk of nb signals (cached on the device)."""
        hit = self._segs.get(nb)
        if hit is not None:
            return hit
        rows, line, nbytes = [], 0, 0.0

        def add(which, off, ss_g, ss_part, nparts, level, N, chan_dev, NI):
            nonlocal line, nbytes
            rows.append((off, ss_g, ss_part, self.phi_lv[level].data_ptr() if self.phi_lv is not None else 0,
                         chan_dev.data_ptr(), which, nparts, N, self.Fc[level], NI, line))
            line += nb * NI
            nbytes += nb * NI * (nparts * self.Fc[level] * 8 + self.M * 4)

        add(0, 0, self.Np, 0, 1, 0, self.Np, self.chan0, 1)
        for g in self.groups:
            NI, N1 = g["NI"], g["N1"]
            if g["children"]:
                add(1, nb * g["u1_off"], N1, 0, 1, g["k1"], N1, g["chan_dev"], NI)
                for c in g["children"]:
                    Fc = self.Fc[c["level"]]
                    add(2, nb * c["part_off"], c["nparts"] * Fc, Fc, c["nparts"], c["level"], c["N2"], c["chan_dev"], NI)
            else:
                Fc = self.Fc[g["k1"]]
                add(2, nb * g["part_off"], g["nparts"] * Fc, Fc, g["nparts"], g["k1"], N1, g["chan_dev"], NI)
        arr = np.array(rows, dtype=_FINSEG)
        dev = torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)
        hit = self._segs[nb] = (dev, len(rows), line, nbytes)
        return hit

    # ------------------------------------------------------------------------------------------------
    def chunk_size(self, B):
        budget = int(os.environ.get("SCAT_B200_WS1D_MB", "6144")) << 20
        per = max(1, 8 * sum(self.per_signal.values()))
        return max(1, min(B, budget // per))

    def rfft(self, U0):
        """U0: (B, Np) float32 padded real signals -> (B, Np, 2) natural-order spectrum (core/scattering1d.py:41)."""
        B = U0.shape[0]
        out = torch.empty((B, self.Np, 2), dtype=torch.float32, device=self.device)
        if B:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.scat1d_rfft(self.tables.for_length(self.Np).data_ptr(), U0.data_ptr(), out.data_ptr(),
                                                out.data_ptr(), B, self.Np, _stream(self.device)))
        return out

    def forward_unaveraged(self, U0_hat):
        """T=0: -> (mods1, mods2): mods1[gi] is the (B, NI, N1) float32 modulus of first-order group gi in natural time
        order, mods2[(gi, ci)] the (B, NI, N2) modulus of its ci-th second-order child group."""
        lib, dev = self.lib, self.device
        B, Np = U0_hat.shape[0], self.Np
        mods1 = [torch.empty((B, g["NI"], g["N1"]), dtype=torch.float32, device=dev) for g in self.groups]
        mods2 = {(gi, ci): torch.empty((B, g["NI"], c["N2"]), dtype=torch.float32, device=dev)
                 for gi, g in enumerate(self.groups) for ci, c in enumerate(g["children"])}
        if B == 0:
            return mods1, mods2
        Bc = self.chunk_size(B)
        ps = self.per_signal
        with torch.cuda.device(dev):
            st = _stream(dev)
            Y = torch.empty(max(1, Bc * ps["Y"]) * 8, dtype=torch.uint8, device=dev)
            U1 = torch.empty(max(1, Bc * ps["U1"]) * 8, dtype=torch.uint8, device=dev)
            yp, up = Y.data_ptr(), U1.data_ptr()
            for b0 in range(0, B, Bc):
                nb = min(Bc, B - b0)
                u0 = U0_hat.data_ptr() + b0 * Np * 8
                for gi, g in enumerate(self.groups):
                    NI, N1, G = g["NI"], g["N1"], nb * g["NI"]
                    tab = g["tab"].data_ptr()
                    leaf = not g["children"]
                    u1 = up + nb * g.get("u1_off", 0) * 8
                    m1 = mods1[gi].data_ptr() + b0 * NI * N1 * 4
                    rd1 = float(nb) * 8 * sum(g["supp_len"])
                    if g["tile"]:
                        _lib.check(lib.scat1d_tile_t0(tab, u0, Np, 0, g["filt_dev"].data_ptr(), g["supp_dev"].data_ptr(),
                                                      None if leaf else u1, m1, G, NI, Np, N1,
                                                      rd1 + float(G) * N1 * (4 if leaf else 12), st))
                    else:
                        _lib.check(lib.scat1d_col_prod(tab, u0, Np, 0, g["filt_dev"].data_ptr(), g["supp_dev"].data_ptr(),
                                                       yp, G, NI, Np, N1, rd1 + float(G) * 8 * N1, st))
                        _lib.check(lib.scat1d_row_mod_t0(tab, yp, G, N1, m1, int(leaf), float(G) * N1 * (12 if leaf else 20), st))
                        if not leaf:
                            _lib.check(lib.scat1d_col_fwd(tab, yp, u1, G, N1, float(G) * N1 * 16, st))
                    for ci, c in enumerate(g["children"]):
                        N2, ctab = c["N2"], c["tab"].data_ptr()
                        m2 = mods2[(gi, ci)].data_ptr() + b0 * NI * N2 * 4
                        rd2 = float(nb) * 8 * sum(c["supp_len"])
                        if c["tile"]:
                            _lib.check(lib.scat1d_tile_t0(ctab, u1, NI * N1, N1, c["filt_dev"].data_ptr(),
                                                          c["supp_dev"].data_ptr(), None, m2, G, NI, N1, N2,
                                                          rd2 + float(G) * 4 * N2, st))
                        else:
                            _lib.check(lib.scat1d_col_prod(ctab, u1, NI * N1, N1, c["filt_dev"].data_ptr(),
                                                           c["supp_dev"].data_ptr(), yp, G, NI, N1, N2,
                                                           rd2 + float(G) * 8 * N2, st))
                            _lib.check(lib.scat1d_row_mod_t0(ctab, yp, G, N2, m2, 1, float(G) * N2 * 12, st))
        return mods1, mods2

    def forward(self, U0_hat):
        """U0_hat: (B, Np, 2) float32 natural-order spectrum of the padded signals -> (B, K, M) float32:
        every channel's low-passed, subsampled (stride 2^log2_stride) signal BEFORE unpadding."""
        lib, dev = self.lib, self.device
        B, Np, M, K = U0_hat.shape[0], self.Np, self.M, self.K
        out = torch.empty((B, K, M), dtype=torch.float32, device=dev)
        if B == 0:
            return out
        Bc = self.chunk_size(B)
        ps = self.per_signal
        with torch.cuda.device(dev):
            st = _stream(dev)
            Y = torch.empty(max(1, Bc * ps["Y"]) * 8, dtype=torch.uint8, device=dev)
            U1 = torch.empty(max(1, Bc * ps["U1"]) * 8, dtype=torch.uint8, device=dev)
            part = torch.empty(max(1, Bc * ps["part"]) * 8, dtype=torch.uint8, device=dev)
            yp, up, pp = Y.data_ptr(), U1.data_ptr(), part.data_ptr()
            for b0 in range(0, B, Bc):
                nb = min(Bc, B - b0)
                u0 = U0_hat.data_ptr() + b0 * Np * 8
                for g in self.groups:
                    NI, N1, G = g["NI"], g["N1"], nb * g["NI"]
                    tab = g["tab"].data_ptr()
                    rd1 = float(nb) * 8 * sum(g["supp_len"])           # algorithmic reads of the parent spectrum
                    u1 = up + nb * g.get("u1_off", 0) * 8
                    if g["tile"]:
                        Fc = self.Fc[g["k1"]]
                        leaf = not g["children"]
                        _lib.check(lib.scat1d_tile(tab, u0, Np, 0, g["filt_dev"].data_ptr(), g["supp_dev"].data_ptr(),
                                                   None if leaf else u1, pp + nb * g["part_off"] * 8 if leaf else None, Fc,
                                                   G, NI, Np, N1, rd1 + float(G) * 8 * (Fc if leaf else N1), st))
                    else:
                        _lib.check(lib.scat1d_col_prod(tab, u0, Np, 0, g["filt_dev"].data_ptr(), g["supp_dev"].data_ptr(),
                                                       yp, G, NI, Np, N1, rd1 + float(G) * 8 * N1, st))
                        if g["children"]:
                            _lib.check(lib.scat1d_row_mod(tab, yp, G, N1, None, 0, float(G) * N1 * 16, st))
                            _lib.check(lib.scat1d_col_fwd(tab, yp, u1, G, N1, float(G) * N1 * 16, st))
                        else:
                            Fc = self.Fc[g["k1"]]
                            _lib.check(lib.scat1d_row_mod(tab, yp, G, N1, pp + nb * g["part_off"] * 8, Fc,
                                                          float(G) * 8 * (N1 + g["nparts"] * Fc), st))
                    for c in g["children"]:
                        N2, ctab = c["N2"], c["tab"].data_ptr()
                        Fc = self.Fc[c["level"]]
                        rd2 = float(nb) * 8 * sum(c["supp_len"])
                        cpart = pp + nb * c["part_off"] * 8
                        if c["tile"]:
                            _lib.check(lib.scat1d_tile(ctab, u1, NI * N1, N1, c["filt_dev"].data_ptr(),
                                                       c["supp_dev"].data_ptr(), None, cpart, Fc, G, NI, N1, N2,
                                                       rd2 + float(G) * 8 * Fc, st))
                        else:
                            _lib.check(lib.scat1d_col_prod(ctab, u1, NI * N1, N1, c["filt_dev"].data_ptr(),
                                                           c["supp_dev"].data_ptr(), yp, G, NI, N1, N2,
                                                           rd2 + float(G) * 8 * N2, st))
                            _lib.check(lib.scat1d_row_mod(ctab, yp, G, N2, cpart, Fc,
                                                          float(G) * 8 * (N2 + c["nparts"] * Fc), st))
                segs, nseg, lines, nbytes = self._segments(nb)
                if self.average_global:
                    _lib.check(lib.scat1d_finish_global(u0, up, pp, segs.data_ptr(), nseg, lines,
                                                        out.data_ptr() + b0 * K * 4, K, st))
                else:
                    _lib.check(lib.scat1d_finish(self.fin_tab.data_ptr(), u0, up, pp, segs.data_ptr(), nseg, lines, M,
                                                 out.data_ptr() + b0 * K * M * 4, K * M, 0, M, nbytes, st))
        return out
