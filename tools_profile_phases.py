"""Phase breakdown of the persistent kernel (in-kernel cycle counters): python tools_profile_phases.py [workload] [steps]"""
import json
import sys

import numpy as np

import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[wl]
X, y, prm = bench.make_data(w)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
s = PySampler.init(X, y, DeviceLikelihood(w["family"], 1.0), st)
for _ in range(2):
    s.step_device(False)
s.sync()
p0 = s.profile()
s.timer_start()
for _ in range(steps):
    s.step_device(False)
ms = s.timer_stop()
s.sync()
p1 = s.profile()
ghz = 1.965
out = {"workload": wl, "ms_per_sweep": ms / steps}
for k in ("pass_cycles", "wait_serial_cycles", "serial_cycles", "passes"):
    out[k] = {n: (p1[k][n] - p0[k][n]) / steps for n in p1[k]}
print(json.dumps(out))
print(f"{'phase':8s} {'passes/sweep':>12s} {'pass us':>10s} {'wait+serial us':>15s} {'serial us':>10s}   (per pass, at {ghz} GHz)")
tot = 0.0
for n in ("root", "minmax", "stats", "bern", "part", "final"):
    cnt = out["passes"][n]
    if cnt == 0:
        continue
    a, b, c = (out[k][n] / cnt / ghz / 1e3 for k in ("pass_cycles", "wait_serial_cycles", "serial_cycles"))
    tot += (out["pass_cycles"][n] + out["wait_serial_cycles"][n]) / ghz / 1e6
    print(f"{n:8s} {cnt:12.1f} {a:10.2f} {b:15.2f} {c:10.2f}")
for n in p1["serial_sub_cycles"]:
    cyc = (p1["serial_sub_cycles"][n] - p0["serial_sub_cycles"][n]) / steps
    calls = (p1["serial_sub_calls"][n] - p0["serial_sub_calls"][n]) / steps
    if calls:
        print(f"  serial {n:14s} calls/sweep {calls:7.1f}  us/call {cyc / calls / ghz / 1e3:7.2f}  ms/sweep {cyc / ghz / 1e6:6.2f}")
c = s.counters()
print("counters", c, "speculative root passes per sweep", (p1["speculative_root_passes"] - p0["speculative_root_passes"]) / steps)
print(f"sum of block-0 pass+wait time: {tot:.2f} ms per sweep; measured {ms / steps:.2f} ms")
