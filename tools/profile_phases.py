"""Phase breakdown of the persistent kernel (in-kernel cycle counters): python tools/profile_phases.py [workload] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

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
if len(sys.argv) > 3:
    s.set_option("root_pass", sys.argv[3])
if len(sys.argv) > 4:
    s.set_option("prefetch_next", sys.argv[4])
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
ctl = {k: (p1["control"][k] - p0["control"][k]) / steps for k in p1["control"]}
us = lambda cyc: cyc / ghz / 1e3
print(f"control block: root sections with speculative issue {ctl['root_spec_n']:.1f}/sweep x {us(ctl['root_spec_cycles']) / max(ctl['root_spec_n'], 1):.2f} us; "
      f"without {ctl['root_nospec_n']:.1f}/sweep x {us(ctl['root_nospec_cycles']) / max(ctl['root_nospec_n'], 1):.2f} us")
print("control block idle (ms/sweep) waiting for: " + ", ".join(f"{k[5:]} {us(ctl[k]) / 1e3:.2f}" for k in ("wait_root", "wait_stats", "wait_part", "wait_other")))
for n in ("root", "stats", "part"):
    d = {k: (p1["diag"][n][k] - p0["diag"][n][k]) / steps for k in p1["diag"][n]}
    cnt = out["passes"][n]
    if cnt and d["issue_to_completion_ns"]:
        print(f"diag {n}: issue->last block saw command {d['issue_to_last_seen_ns'] / cnt / 1e3:.2f} us, issue->last arrival "
              f"{d['issue_to_last_arrival_ns'] / cnt / 1e3:.2f} us, issue->completion seen {d['issue_to_completion_ns'] / cnt / 1e3:.2f} us, "
              f"longest pass body {d['longest_body_cycles'] / cnt / ghz / 1e3:.2f} us")
bc = s.block_cycles()
if bc.any():
    for ph, n in ((1, "root"), (3, "stats"), (4, "bern"), (5, "part")):
        v = bc[:, ph] / ghz / 1e3
        order = np.argsort(v)
        print(f"per-block total {n} body time (us): min {v.min():.0f} median {np.median(v):.0f} max {v.max():.0f}; slowest blocks {order[-6:].tolist()} "
              f"-> {np.round(v[order[-6:]]).tolist()}; fastest {order[:4].tolist()} -> {np.round(v[order[:4]]).tolist()}")
if bc.any() and bc.shape[1] >= 16:
    nroot = out["passes"]["root"] * (steps + 2)
    for b in (0, 73, 146):
        print(f"staged root pass checkpoints, block {b} (us/pass): " + ", ".join(f"{n} {bc[b, 8 + i] / nroot / ghz / 1e3:.2f}" for i, n in enumerate(["prologue + (in loop) A/y->o incl. tree eval", "prologue + (in loop) A store", "issue first tiles (+ in-loop issue)", "first tile lands + (in loop) totals+jobs", "tile loop", "epilogue", "(in loop) wait own copies", "(in loop) block sync"])))
