"""Long free-running steps on the generic control code (P > 32, weight_mode pg_correct, OneHotSplit columns).  python tools/soak_generic.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
from tests import datagen

def run(name, X, y, m, P, fam, alpha, beta, md, steps, rules=None, mode=None):
    p = X.shape[1]
    prm = datagen.pgbart_params(y, m, p, alpha, beta)
    st = PyBartSettings(m, P, prm["max_depth"] if md is None else md, alpha, beta, prm["sigma"], list(prm["split_prior"]),
                        rules or ["ContinuousSplit"] * p, "constant", "systematic", 1.0, 1.0, seed=7)
    s = PySampler.init(X, y, DeviceLikelihood(fam, 1.0), st)
    if mode:
        s.set_option("weight_mode", mode)
    t0 = time.perf_counter()
    for i in range(steps):
        s.step_device(i < steps // 2)
        if i % 32 == 31:
            s.sync()
    s.sync()
    dt = time.perf_counter() - t0
    pred = s.predictions()
    assert np.all(np.isfinite(pred))
    depth = s.info()[2]
    print(f"{name}: {steps} steps in {dt:.1f} s ({1e3 * dt / steps:.2f} ms/step), {int(s.counters()['tree_updates'])} tree updates, "
          f"max tree depth of the last step {int(depth.max()) if depth.size else 0}, rmse {np.sqrt(np.mean((pred - y) ** 2)):.3f}", flush=True)
    s.close()

X, y = datagen.friedman(50_000, 10, seed=3, bernoulli=True)
run("P=64 bernoulli", X, y, 40, 64, "bernoulli_logit", 0.95, 2.0, None, 150)
X, y = datagen.friedman(50_000, 10, seed=4)
run("P=256 gaussian", X, y, 20, 256, "gaussian", 0.95, 2.0, None, 100)
run("pg_correct P=16", X[:20_000], y[:20_000], 30, 16, "gaussian", 0.95, 1.5, 12, 400, mode="pg_correct")
rng = np.random.default_rng(0)
Xc = X[:30_000].copy(); Xc[:, 0] = rng.integers(0, 5, 30_000); Xc[:, 1] = rng.integers(-2, 3, 30_000)
yc = (y[:30_000] + Xc[:, 0]).astype(np.float32).astype(np.float64)
run("onehot + pg_correct P=12", Xc, yc, 30, 12, "gaussian", 0.95, 1.5, 12, 300, rules=["OneHotSplit", "OneHotSplit"] + ["ContinuousSplit"] * 8, mode="pg_correct")
run("onehot reference mode P=20", Xc, yc, 50, 20, "gaussian", 0.95, 2.0, None, 400, rules=["OneHotSplit", "OneHotSplit"] + ["ContinuousSplit"] * 8)
print("generic soak ok")
