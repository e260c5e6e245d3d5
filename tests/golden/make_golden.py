"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference or baseline/_ref):

    python tests/golden/make_golden.py

Outputs (all small .npz, committed):
  ref_fixture_{1,2,3}d.npz   verbatim re-save of the reference's own
                             tests/scattering{1,2,3}d/test_data_{1,2,3}d.npz
  golden_2d_*.npz            reference numpy frontend on seeded inputs
                             (float64 input -> float64 oracle)
  golden_filters_2d.npz      checksums + samples of the reference filter bank
  golden_filters_3d.npz      reference solid-harmonic / Gaussian banks (small: full arrays; 32^3: checksums + samples)
  golden_1d_*.npz, golden_3d_*.npz  the same for the 1D / 3D frontends

The reference is imported with the sph_harm shim (scipy >= 1.15 removed
scipy.special.sph_harm, used at kymatio/scattering3d/filter_bank.py:4,167).
"""
import os
import sys

import numpy as np
import scipy.special

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isdir(os.path.join(cand, "kymatio")):
        sys.path.insert(0, cand)
        REF_ROOT = cand
        break
else:
    raise SystemExit("reference not found")

if not hasattr(scipy.special, "sph_harm"):
    scipy.special.sph_harm = lambda m, n, az, pol: scipy.special.sph_harm_y(n, m, pol, az)

from kymatio.numpy import Scattering1D, Scattering2D, HarmonicScattering3D  # noqa: E402
from kymatio.scattering2d.filter_bank import filter_bank as ref_filter_bank_2d  # noqa: E402


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print("wrote", name, {k: np.asarray(v).shape for k, v in kw.items()})


def resave_fixtures():
    src = "/root/reference/tests"
    if not os.path.isdir(src):
        print("reference tests not mounted; skipping fixture re-save")
        return
    for d in (1, 2, 3):
        data = np.load(os.path.join(src, f"scattering{d}d", f"test_data_{d}d.npz"), allow_pickle=True)
        save(f"ref_fixture_{d}d.npz", **{k: data[k] for k in data.files})


def golden_2d():
    cases = [
        # name, J, L, shape, batch, max_order, pre_pad
        ("c1_J2_32", 2, 8, (32, 32), 4, 2, False),        # BASELINE configs[0] shape
        ("c2_J3_256", 3, 8, (256, 256), 2, 2, False),      # headline shape (reduced batch)
        ("c5_J4_224", 4, 8, (224, 224), 1, 2, False),      # configs[4] shape
        ("J3_64", 3, 8, (64, 64), 3, 2, False),
        ("J2_33x47", 2, 8, (33, 47), 2, 2, False),          # odd, non-square
        ("J1_31", 1, 8, (31, 31), 2, 2, False),
        ("J4_32_L4", 4, 4, (32, 32), 2, 2, False),
        ("J5_32", 5, 8, (32, 32), 1, 2, False),            # 1x1 output
        ("J3_64_o1", 3, 8, (64, 64), 2, 1, False),         # max_order=1
        ("J2_prepad48", 2, 8, (32, 32), 2, 2, True),       # pre_pad input 48x48? see below
        ("J3_240", 3, 8, (224, 224), 1, 2, False),         # 15*16 padded size
        ("J2_24x40", 2, 6, (24, 40), 2, 2, False),
    ]
    rng = np.random.RandomState(42)
    for name, J, L, shape, B, mo, pre_pad in cases:
        S = Scattering2D(J, shape, L=L, max_order=mo, pre_pad=pre_pad)
        in_shape = (S._M_padded, S._N_padded) if pre_pad else shape
        x = rng.randn(B, *in_shape)
        Sx64 = S(x)                      # float64 path
        x32 = x.astype(np.float32)
        Sx32 = S(x32)                    # reference fp32 numpy path
        save(f"golden_2d_{name}.npz", x=x32, Sx64=Sx64.astype(np.float64), Sx32=Sx32,
             J=J, L=L, shape=np.array(shape), max_order=mo, pre_pad=pre_pad)


def golden_filters_2d():
    out = {}
    for (Mp, Np, J, L) in [(40, 40, 2, 8), (272, 272, 3, 8), (48, 64, 2, 6), (64, 64, 4, 8)]:
        fb = ref_filter_bank_2d(Mp, Np, J, L)
        key = f"{Mp}x{Np}_J{J}_L{L}"
        sums, l2, samples = [], [], []
        for lev in fb["phi"]["levels"]:
            sums.append(lev.sum(dtype=np.float64)); l2.append(np.sqrt((lev.astype(np.float64) ** 2).sum()))
            samples.append(lev.ravel()[:: max(1, lev.size // 16)][:16].astype(np.float64))
        for p in fb["psi"]:
            for lev in p["levels"]:
                sums.append(lev.sum(dtype=np.float64)); l2.append(np.sqrt((lev.astype(np.float64) ** 2).sum()))
                samples.append(lev.ravel()[:: max(1, lev.size // 16)][:16].astype(np.float64))
        out[key + "_sum"] = np.array(sums)
        out[key + "_l2"] = np.array(l2)
        out[key + "_samples"] = np.stack(samples)
        if Mp <= 64:
            out[key + "_phi0"] = fb["phi"]["levels"][0]
            out[key + "_psi_last_l0"] = fb["psi"][-1]["levels"][0]
            out[key + "_psi3_l0"] = fb["psi"][3]["levels"][0]
    save("golden_filters_2d.npz", **out)


def golden_1d():
    rng = np.random.RandomState(7)
    cases = [
        ("J5_Q4_2048", dict(J=5, shape=2048, Q=(4, 1)), 3),
        ("J8_Q8_65536", dict(J=8, shape=2 ** 16, Q=(8, 1)), 1),   # configs[2] shape, batch 1
        ("J6_Q16_512", dict(J=6, shape=512, Q=16), 2),
        ("J4_Q2_1000_o1", dict(J=4, shape=1000, Q=2, max_order=1), 2),
    ]
    for name, kw, B in cases:
        S = Scattering1D(**kw)
        x = rng.randn(B, kw["shape"])
        Sx = S(x)
        meta = S.meta()
        save(f"golden_1d_{name}.npz", x=x.astype(np.float32), Sx64=np.asarray(Sx, dtype=np.float64),
             order=meta["order"], key=np.array([str(k) for k in meta["key"]]),
             **{k: np.array(v) for k, v in kw.items()})


def golden_3d_c4():
    """BASELINE configs[3] at its own size (J=2, L=2, 128^3): the 8 MB input is NOT stored - it is regenerated from the
    seed (legacy RandomState streams are stable across numpy versions); only the (1, 6, 3, 3) output is committed."""
    x = np.random.RandomState(1234).randn(1, 128, 128, 128).astype(np.float32)
    S = HarmonicScattering3D(J=2, shape=(128, 128, 128), L=2)
    Sx = S(x.astype(np.float64))
    save("golden_3d_c4_J2_L2_128.npz", seed=1234, Sx64=np.asarray(Sx, dtype=np.float64), J=2, L=2,
         shape=np.array((128, 128, 128)), x_first8=x.ravel()[:8], x_sum=np.float64(x.astype(np.float64).sum()))


def golden_filters_3d():
    """Reference solid-harmonic / Gaussian banks: full arrays at small (even and odd) sizes, checksums at 32^3."""
    from kymatio.scattering3d.filter_bank import solid_harmonic_filter_bank, gaussian_filter_bank
    out = {}
    for (M, N, O, J, L, s0) in [(8, 10, 12, 2, 3, 1.0), (9, 8, 7, 1, 2, 1.5), (32, 32, 32, 2, 2, 1.0)]:
        key = f"{M}x{N}x{O}_J{J}_L{L}"
        bank = solid_harmonic_filter_bank(M, N, O, J, L, s0)
        gauss = gaussian_filter_bank(M, N, O, J, s0)
        out[key + "_cfg"] = np.array([M, N, O, J, L, s0], dtype=np.float64)
        if M * N * O <= 1000:
            for l, b in enumerate(bank):
                out[key + f"_l{l}"] = b
            out[key + "_gauss"] = gauss
        else:
            for l, b in enumerate(bank):
                out[key + f"_l{l}_sum"] = b.reshape(b.shape[0], b.shape[1], -1).astype(np.complex128).sum(-1)
                out[key + f"_l{l}_l2"] = np.sqrt((np.abs(b.reshape(b.shape[0], b.shape[1], -1).astype(np.complex128)) ** 2).sum(-1))
                out[key + f"_l{l}_samples"] = b.reshape(b.shape[0], b.shape[1], -1)[:, :, ::997]
            out[key + "_gauss_samples"] = gauss.reshape(gauss.shape[0], -1)[:, ::997]
    save("golden_filters_3d.npz", **out)


def golden_3d():
    rng = np.random.RandomState(11)
    cases = [
        ("J2_L2_16", dict(J=2, shape=(16, 16, 16), L=2), 2),
        ("J2_L2_32", dict(J=2, shape=(32, 32, 32), L=2), 1),
        ("J1_L3_12x16x20", dict(J=1, shape=(12, 16, 20), L=3, integral_powers=(1.0, 2.0)), 2),
    ]
    for name, kw, B in cases:
        S = HarmonicScattering3D(**kw)
        x = rng.randn(B, *kw["shape"])
        Sx = S(x)
        kw2 = dict(kw)
        kw2["shape"] = np.array(kw["shape"])
        save(f"golden_3d_{name}.npz", x=x.astype(np.float32), Sx64=np.asarray(Sx, dtype=np.float64),
             **{k: np.array(v) for k, v in kw2.items()})


if __name__ == "__main__":
    resave_fixtures()
    golden_filters_2d()
    golden_filters_3d()
    golden_2d()
    golden_1d()
    golden_3d()
    golden_3d_c4()
