"""Compile one .cu of kymatio_b200/csrc with -Xptxas -v and print registers / spills / smem per kernel.
usage: python tools/ptxas_info.py tile_inst.cu [filter-regex] [extra nvcc flags...]"""
import re
import subprocess
import sys
import os

src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else "."
extra = sys.argv[3:]
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kymatio_b200", "csrc")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr",
       "-diag-suppress", "20013", "-Xptxas", "-v", "-c", os.path.join(csrc, src), "-o", "/tmp/ptxas_info.o"] + extra
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(sb::\w+<\w+>\)", "", name)
        continue
    if "bytes stack frame" in line:
        stack = line.strip().replace("ptxas info    : ", "")
    m = re.search(r"Used (\d+) registers", line)
    if m and name and re.search(flt, name):
        print(f"{name:70s} regs {m.group(1):>3s}  {stack}")
    if "error" in line.lower():
        print(line)
