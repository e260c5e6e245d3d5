"""Aggregate an exported ncu source page (tools/ncu_export.sh -> *_source.csv.gz).
usage: python tools/ncu_src.py file.csv.gz            -> list kernel blocks
       python tools/ncu_src.py file.csv.gz IDX [N]    -> top-N source lines + opcode histogram + stall mix of block IDX"""
import csv, gzip, io, sys, re, collections

path = sys.argv[1]
rd = csv.reader(io.TextIOWrapper(gzip.open(path), newline=""))
blocks, cur, hdr, fpath = [], None, None, None
for r in rd:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]
        if cur is None or cur["fn"] != fn or (cur["files"] and fpath in cur["files"] and cur["files"][-1] != fpath):
            cur = {"fn": fn, "files": [], "rows": []}; blocks.append(cur)
        cur["files"].append(fpath); continue
    if r[0] == "Line No":
        hdr = r; continue
    cur["rows"].append((fpath, r))
if len(sys.argv) < 3:
    for i, b in enumerate(blocks):
        print(i, b["fn"][:90], sorted(set(b["files"])))
    sys.exit()
def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


b = blocks[int(sys.argv[2])]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
iI, iS = ix["Instructions Executed"], ix["# Samples"]
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines, ops, stalls = {}, collections.Counter(), collections.Counter()
seen = set()
for f, r in b["rows"]:
    if r[0].isdigit():
        lines[(f, int(r[0]))] = [num(r[iI]), num(r[iS]), r[1][:95], collections.Counter({h: num(r[i]) for h, i in stall_cols if num(r[i])})]
    elif r[0] == "" and len(r) > iI:
        key = (r[2], r[3])
        if key in seen:
            continue
        seen.add(key)
        op = r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]
        ops[op.split(".")[0]] += num(r[iI])
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print(b["fn"][:100], "total warp-inst", tot, "samples", ts)
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in lines.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
print({f: (round(100 * a / tot, 1), round(100 * s / ts, 1)) for f, (a, s) in byfile.items()})
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:N]:
    top = ",".join(f"{k[6:]}:{c}" for k, c in v[3].most_common(3))
    print(f"{f:14s}:{l:4d} inst {100*v[0]/tot:5.1f}% samp {100*v[1]/ts:5.1f}% [{top}] {v[2]}")
allst = collections.Counter()
for v in lines.values():
    allst.update(v[3])
print("stall mix:", {k[6:]: round(100 * c / max(1, sum(allst.values())), 1) for k, c in allst.most_common(10)})
oi = sum(ops.values())
print("opcodes:", {k: round(100 * c / oi, 1) for k, c in ops.most_common(25)})
