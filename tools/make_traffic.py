"""profiles/traffic.json from an ncu capture of one C2 forward (the table bench.py reads for dram_frac / issue_frac /
l1tex_frac).  On the GPU box:
    python tools/launch_labels.py 256
    ncu --set full --clock-control none -k regex:k2d_ -s <launches of the warm-up forwards> -c <launches per forward> \
        -f -o gpurun_out/c2_full python tools/run_once.py 256 3
    ncu -i gpurun_out/c2_full.ncu-rep --page raw --csv > gpurun_out/c2_raw.csv        (tools/ncu_export.sh does this)
here:
    python tools/make_traffic.py gpurun_out/c2_raw.csv gpurun_out/launch_labels.json profiles/traffic.json"""
import csv, json, sys
raw, labels, out = sys.argv[1], json.load(open(sys.argv[2])), sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(labels["labels"]), (len(data), len(labels["labels"]))


def f(r, k):
    return float(r[ix[k]].replace(",", ""))


unit_r, unit_w = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
table = {}
for lab, r in zip(labels["labels"], data):
    table[lab] = {
        "batch": labels["batch"], "kernel": r[ix["Kernel Name"]],
        "dram_bytes_per_launch": f(r, "dram__bytes_read.sum") * scale[unit_r] + f(r, "dram__bytes_write.sum") * scale[unit_w],
        "issue_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1tex_pct": f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "lts_pct": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "ncu_duration_ms": f(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(rows[1][ix["gpu__time_duration.sum"]], 1.0),
        "registers": int(f(r, "launch__registers_per_thread")),
    }
json.dump(table, open(out, "w"), indent=1)
print("wrote", out, len(table), "kernels")
