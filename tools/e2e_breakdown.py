"""Where the end-to-end step time goes beyond the sweep itself: python tools/e2e_breakdown.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
w = bench.WORKLOADS[wl]
X, y, prm = bench.make_data(w)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
lik = DeviceLikelihood(w["family"], 1.0)
s = PySampler.init(X, y, lik, st)
for _ in range(3):
    s.step(False)
K = 10
def t(fn):
    t0 = time.perf_counter(); fn(); return (time.perf_counter() - t0) * 1e3 / K
a = t(lambda: [ (s.step_device(False), s.sync()) for _ in range(K)])
b = t(lambda: [ s.step(False) for _ in range(K)])
c = t(lambda: [ (s.set_likelihood(lik), s.step(False)) for _ in range(K)])
s.timer_start()
for _ in range(K): s.step_device(False)
dev = s.timer_stop() / K
print(f"{wl}: device-timed sweep {dev:.3f} ms; step_device+sync {a:.3f} ms; step() with D2H {b:.3f} ms; set_likelihood+step() {c:.3f} ms")
# raw pinned D2H of the same size
n = X.shape[0]
d = torch.empty(n, dtype=torch.float64, device="cuda"); h = torch.empty(n, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for _ in range(3): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"pinned D2H of {n * 8 / 1e6:.1f} MB: {dt * 1e3:.3f} ms = {n * 8 / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(200): torch.cuda.synchronize()
print(f"empty synchronize: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
