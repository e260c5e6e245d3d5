"""Summarise an .ncu-rep: per-kernel key metrics, and (optionally) per-source-line instruction shares.
usage: python tools/ncu_summary.py rep.ncu-rep [--lines KERNEL_REGEX LAUNCH_INDEX]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
for r in data:
    print("----", r[ix["Kernel Name"]][:70], r[ix["Grid Size"]], r[ix["Block Size"]])
    for w in want:
        if w in ix:
            print(f"   {w:66s} {r[ix[w]]:>16s} {units[ix[w]]}")
if "--lines" in sys.argv:
    i = sys.argv.index("--lines")
    kre, li = sys.argv[i + 1], sys.argv[i + 2]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kre, "--launch-skip", li, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    cur, agg, h = None, {}, None
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r and r[0] == "Line No":
            h = r; continue
        if h and len(r) == len(h) and r[0].isdigit():
            a = agg.setdefault((cur, int(r[0])), [0, 0, r[1][:100]])
            a[0] += int(r[7]); a[1] += int(r[6])
    tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
    print("total warp-inst", tot, "samples", ts)
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{f:14s}:{l:4d} inst {100*a[0]/tot:5.1f}% samp {100*a[1]/ts:5.1f}%  {a[2]}")
