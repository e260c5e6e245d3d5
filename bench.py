#!/usr/bin/env python
"""bench.py - the BASELINE.json headline: Scattering2D J=3 L=8 256x256 fp32 images/s.

    python bench.py --gpus N --steps K --warmup W            # this repo (torch_b200)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one forward of the scattering hot path over one batch of synthetic images
(configs[1]: batch 256 per GPU, fixed as N grows -> weak scaling).  Rank 0 prints ONE
JSON line.  See DESIGN.md "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

J, L, SHAPE = 3, 8, (256, 256)
PASS_MODEL_BYTES_PER_IMAGE = 107_096_064      # SURVEY 8(d), per-(j,theta)-pass model, C2
METRIC = "scattering2d_J3_L8_256x256_images_per_s"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=96)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# reference (CPU) arm helpers
# ------------------------------------------------------------------------------------------
def _import_reference():
    """The unmodified reference from baseline/_ref (pip --target install)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "kymatio")):
        return None
    import scipy.special
    if not hasattr(scipy.special, "sph_harm"):   # kymatio/scattering3d/filter_bank.py:4 vs scipy >= 1.15
        scipy.special.sph_harm = lambda m, n, az, pol: scipy.special.sph_harm_y(n, m, pol, az)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from kymatio.numpy import Scattering2D
    return Scattering2D


_REF_S = None


def _ref_worker(x):
    return _REF_S(x).shape[0]


def _make_cpu_scattering():
    """-> (callable(batch ndarray) -> coefficients, kind)"""
    RefScattering2D = _import_reference()
    if RefScattering2D is not None:
        return RefScattering2D(J, SHAPE, L=L), "reference"
    from oracle import scattering2d as o2
    Mp, Np = o2.padded_size(SHAPE[0], SHAPE[1], J)
    fb = o2.filter_bank(Mp, Np, J, L)
    return (lambda x: o2.scattering2d(x, J, L, filters=fb)), "port"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_single(sample):
    import numpy as np
    S, kind = _make_cpu_scattering()
    x = np.random.RandomState(42).randn(sample, *SHAPE).astype("float32")
    S(x[:2])
    t0 = time.perf_counter()
    for i in range(0, sample, 8):
        S(x[i:i + 8])
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{sample} images 256x256 fp32, numpy backend, 1 process, {dt:.1f} s"}


def reference_torch_gpu_baseline(dev, batch=64, steps=3):
    """The reference's own torch GPU backend (cuFFT + ATen, kymatio/scattering2d/backend/torch_backend.py) on the same
    device and config, reduced batch - a reported baseline beside the numpy CPU one (north_star); None if the reference
    is not installed under baseline/_ref."""
    if _import_reference() is None:
        return None
    import torch
    try:
        from kymatio.scattering2d.frontend.torch_frontend import ScatteringTorch2D
        S = ScatteringTorch2D(J, SHAPE, L=L, backend="torch").to(dev)
        x = torch.randn(batch, *SHAPE, dtype=torch.float32, device=dev)
        with torch.no_grad():
            S(x)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                S(x)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / ms * 1e3, "unit": UNIT, "batch": batch, "ms_per_step": ms,
                "what": "unmodified kymatio Scattering2D(backend='torch') on the same B200, inputs resident in HBM"}
    except Exception as e:                                     # a baseline must never break the measured arm
        return {"unavailable": repr(e)[:200]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    global _REF_S
    _REF_S, kind = _make_cpu_scattering()
    cores = host_cores()
    per_worker = 2
    sample = cores * per_worker
    x = np.random.RandomState(42).randn(sample, *SHAPE).astype("float32")
    chunks = [x[i * per_worker:(i + 1) * per_worker] for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_worker, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, chunks)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} images/step ({per_worker} per process x {cores} processes), numpy backend"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Scattering2D J=3 L=8 shape=(256,256) fp32 (BASELINE configs[1]), CPU sample",
                   "images_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        in_window = [(t, l) for t, l in self.rows if t_begin <= t <= t_end + 0.1]
        # a very short timed region can fall between two samples: then use the samples taken since the
        # sampler started (warm-up included, the GPU is under the same load)
        for t, line in (in_window or self.rows):
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except Exception:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# torch_b200 arm
# ------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from kymatio_b200 import Scattering2D, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (torch_b200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    S = Scattering2D(J, SHAPE, L=L).to(dev)
    torch.manual_seed(42 + rank)
    x = torch.randn(B, *SHAPE, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    sampler = ClockSampler(local)
    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            y = S(x)
        barrier()

        # ---- device-resident timing: K steps, L2 flushed between steps (not timed) ---------------
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        n0 = _lib.launch_count()
        t_begin = time.perf_counter()
        barrier()
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            y = S(x)
            e1.record()
        barrier()
        t_end = time.perf_counter()
        launches = _lib.launch_count() - n0
        clocks = sampler.stop(t_begin, t_end)
        dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)

        # ---- end to end: pinned host in -> H2D -> forward -> D2H pinned host out, every step -------
        xh = torch.randn(B, *SHAPE, dtype=torch.float32).pin_memory()
        yh = torch.empty((B,) + tuple(y.shape[1:]), dtype=torch.float32).pin_memory()
        nchunk = int(os.environ.get("BENCH_E2E_CHUNKS", "8"))
        if B % nchunk or B < 4 * nchunk:
            nchunk = 1
        cb = B // nchunk
        streams = [torch.cuda.Stream(device=dev) for _ in range(min(int(os.environ.get("BENCH_E2E_STREAMS", "3")), nchunk))]

        def e2e_step():
            for c in range(nchunk):
                st = streams[c % len(streams)]
                with torch.cuda.stream(st):
                    xd = xh[c * cb:(c + 1) * cb].to(dev, non_blocking=True)
                    yd = S(xd)
                    yh[c * cb:(c + 1) * cb].copy_(yd, non_blocking=True)
            for st in streams:
                st.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()

        # ---- live per-kernel timing for the roofline of the dominant kernel -----------------------
        kern = None
        if rank == 0:
            _lib.timing_enable(True)
            for _ in range(2):
                flush.zero_()
                S(x)
            rows = _lib.timing_report()
            _lib.timing_enable(False)
            tot = sum(r["ms"] for r in rows)
            rows.sort(key=lambda r: -r["ms"])
            kern = {"rows": rows, "total_ms": tot}

    # max over ranks (device time and e2e wall time)
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    value = world * B * args.steps / (dev_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    top = kern["rows"][0]
    achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (same batch only)
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top["label"])
        if t and int(t["batch"]) == B:
            traffic = float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Scattering2D J=3 L=8 shape=(256,256) fp32 forward (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-sharded x{world}",
                   "l2": "flushed between timed steps (256 MiB write, untimed); step working set 3.6 GB >> L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": xh.numel() * 4,
                "d2h_bytes_per_step": yh.numel() * 4,
                "note": f"pinned host in/out, {nchunk} chunks pipelined on {len(streams)} streams, wall clock, max over ranks"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "kernel": top["label"],
                     "kernel_share_of_step": top["ms"] / kern["total_ms"], "peak_source": peak_src,
                     "step_pass_model": {"bytes_per_image": PASS_MODEL_BYTES_PER_IMAGE,
                                         "achieved": value / world * PASS_MODEL_BYTES_PER_IMAGE / 1e9,
                                         "frac": value / world * PASS_MODEL_BYTES_PER_IMAGE / 1e9 / hbm_peak}},
        "kernels": [{"label": r["label"], "ms_per_step": r["ms"] / 2, "launches_per_step": r["count"] // 2,
                     "GBps": r["bytes"] / (r["ms"] * 1e-3) / 1e9} for r in kern["rows"][:12]],
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(args.cpu_sample)
        line["reference_torch_gpu"] = reference_torch_gpu_baseline(dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
