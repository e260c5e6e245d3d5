import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f img/s  ms/step %.2f  e2e %.0f  launches %d  roofline %s frac %.3f  step_frac %.3f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["kernel"],
    d["roofline"]["frac"], d["roofline"]["step_pass_model"]["frac"]))
print("clocks", d.get("clocks"))
for k in d["kernels"]:
    print("  %-32s %7.3f ms  x%d  %7.0f GB/s" % (k["label"], k["ms_per_step"], k["launches_per_step"], k["GBps"]))
