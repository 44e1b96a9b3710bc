"""Long free-running sweeps on the default path (no trace, no oracle): looks for rare hangs / device errors in the
issue-ahead, early-statistics and speculation machinery.  python tools/soak.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
from tests import datagen

CASES = [  # n, p, m, P, family, alpha, beta, max_depth, batch, sweeps
    (1_000_000, 50, 200, 20, "gaussian", 0.95, 2.0, None, 1.0, 600),
    (100_000, 20, 200, 20, "gaussian", 0.95, 2.0, None, 1.0, 2500),
    (5_000, 8, 50, 20, "gaussian", 0.95, 2.0, None, 1.0, 4000),
    (20_000, 6, 30, 3, "gaussian", 0.95, 0.7, 10, 0.5, 3000),        # few particles: deep trees, partitions below the root
    (50_000, 10, 40, 8, "gaussian", 0.9999, 1.5, 8, 1.0, 2000),      # every slot grows every round
    (30_000, 8, 40, 12, "bernoulli_logit", 0.95, 2.0, None, 1.0, 2000),
]
for (n, p, m, P, fam, alpha, beta, md, batch, sweeps) in CASES:
    X, y = datagen.friedman(n, p, seed=n % 97, bernoulli=(fam == "bernoulli_logit"))
    prm = datagen.pgbart_params(y, m, p, alpha, beta)
    st = PyBartSettings(m, P, prm["max_depth"] if md is None else md, alpha, beta, prm["sigma"], list(prm["split_prior"]),
                        ["ContinuousSplit"] * p, "constant", "systematic", batch, batch, seed=n)
    s = PySampler.init(X, y, DeviceLikelihood(fam, 1.0), st)
    t0 = time.perf_counter()
    for i in range(sweeps):
        s.step_device(i < sweeps // 2)
        if i % 64 == 63:
            s.sync()
    s.sync()
    dt = time.perf_counter() - t0
    pred = s.predictions()
    c, sp, es = s.counters(), s.spec_counters(), s.early_stats_counters()
    assert np.all(np.isfinite(pred)), "non-finite predictions"
    print(f"n={n} p={p} m={m} P={P} {fam} alpha={alpha} beta={beta}: {sweeps} steps in {dt:.1f} s ({1e3 * dt / sweeps:.2f} ms/step), "
          f"{int(c['tree_updates'])} tree updates, rmse {np.sqrt(np.mean((pred - y) ** 2)):.3f}, speculation {sp}, early statistics {es}", flush=True)
    s.close()
print("soak ok")
