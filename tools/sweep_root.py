"""Root-pass variants side by side: python tools/sweep_root.py [workload] [steps]
Each configuration creates its own sampler (the staging area is sized at create time: PGBART_SMEM_KB)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w = bench.WORKLOADS[wl]
X, y, prm = bench.make_data(w)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
CONFIGS = [
    dict(name="direct", smem=163, root="direct"),
    dict(name="ring G4", smem=163, root="ring"),
    dict(name="ring G4 nofence", smem=163, root="ring", dbg=1),
    dict(name="ring G4 nojobs", smem=163, root="ring", dbg=2),
    dict(name="ring G4 norowwork", smem=163, root="ring", dbg=4),
    dict(name="ring G4 norowwork nofence", smem=163, root="ring", dbg=5),
    dict(name="ring G8 NST2 norowwork nofence", smem=163, root="ring", groups=8, dbg=5),
    dict(name="ring G2 norowwork nofence", smem=163, root="ring", groups=2, dbg=5),
    dict(name="ring 227KB G8 nofence", smem=227, root="ring", groups=8, dbg=1),
]
ghz = 1.965
ref = None
for cfg in CONFIGS:
    os.environ["PGBART_SMEM_KB"] = str(cfg["smem"])
    s = PySampler.init(X, y, DeviceLikelihood(w["family"], 1.0), st)
    s.set_option("root_pass", cfg["root"])
    s.set_option("root_tile_groups", str(cfg.get("groups", 0)))
    s.set_option("root_stages", str(cfg.get("stages", 0)))
    s.set_option("root_dbg", str(cfg.get("dbg", 0)))
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
    if ref is None:
        ref = pred
    same = bool((pred == ref).all())
    d = lambda k, n: (p1[k][n] - p0[k][n]) / steps
    root_us = d("pass_cycles", "root") / max(d("passes", "root"), 1) / ghz / 1e3
    wait_us = d("wait_serial_cycles", "root") / max(d("passes", "root"), 1) / ghz / 1e3
    ctl = {k: (p1["control"][k] - p0["control"][k]) / steps for k in p1["control"]}
    spec_us = ctl["root_spec_cycles"] / max(ctl["root_spec_n"], 1) / ghz / 1e3
    nospec_us = ctl["root_nospec_cycles"] / max(ctl["root_nospec_n"], 1) / ghz / 1e3
    print(json.dumps({"config": cfg["name"], "ms_per_sweep": round(ms / steps, 3), "root_pass_us": round(root_us, 2),
                      "root_wait_us": round(wait_us, 2), "ctl_root_section_spec_us": round(spec_us, 2),
                      "ctl_root_section_nospec_us": round(nospec_us, 2), "smem": s.counters()["smem_bytes"],
                      "bit_identical_to_first": same}), flush=True)
    s.close()
