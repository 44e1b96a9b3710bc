"""One sweep with the stepwise driver (every pass / serial section is its own kernel): run under
ncu --metrics gpu__time_duration.sum to get isolated per-kernel durations.  usage: tools/stepwise_once.py [workload] [driver]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
driver = sys.argv[2] if len(sys.argv) > 2 else "stepwise"
w = bench.WORKLOADS[wl]
X, y, prm = bench.make_data(w)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 0.1, 0.1, seed=0)
s = PySampler.init(X, y, DeviceLikelihood(w["family"], 1.0), st)
s.set_option("driver", driver)
for _ in range(3):
    s.step_device(False)
s.sync()
print("ok", s.counters())
