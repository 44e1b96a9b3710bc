"""A/B of sampler options on one workload: python tools/ab_options.py C3 5 "issue_ahead=1" "issue_ahead=0" "root_pass=ring,issue_ahead=1"
Every configuration gets a fresh sampler with the same seed; reports ms per sweep, the phase times the last pass block saw,
and whether the predictions after the timed sweeps are bit-identical to the first configuration's."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler

wl, steps = sys.argv[1], int(sys.argv[2])
configs = sys.argv[3:] or [""]
w = bench.WORKLOADS[wl]
X, y, prm = bench.make_data(w)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
ghz, ref = 1.965, None
for cfg in configs:
    s = PySampler.init(X, y, DeviceLikelihood(w["family"], 1.0), st)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        s.set_option(k, v)
    for _ in range(3):
        s.step_device(False)
    s.sync()
    p0 = s.profile()
    s.timer_start()
    for _ in range(steps):
        s.step_device(False)
    ms = s.timer_stop()
    s.sync()
    p1 = s.profile()
    pred = s.predictions()
    ref = pred if ref is None else ref
    out = {"config": cfg or "(defaults)", "ms_per_sweep": round(ms / steps, 3), "bit_identical_to_first": bool((pred == ref).all()),
           "speculation": s.spec_counters()}
    for n in ("root", "stats", "part"):
        cnt = (p1["passes"][n] - p0["passes"][n]) / steps
        if cnt:
            out[n] = {"per_sweep": round(cnt, 1),
                      "pass_us": round((p1["pass_cycles"][n] - p0["pass_cycles"][n]) / steps / cnt / ghz / 1e3, 2),
                      "idle_before_us": round((p1["wait_serial_cycles"][n] - p0["wait_serial_cycles"][n]) / steps / cnt / ghz / 1e3, 2)}
    ctl = {k: (p1["control"][k] - p0["control"][k]) / steps for k in p1["control"]}
    out["control_idle_ms"] = round(sum(ctl[k] for k in ("wait_root", "wait_stats", "wait_part", "wait_other")) / ghz / 1e6, 3)
    print(json.dumps(out), flush=True)
    s.close()
