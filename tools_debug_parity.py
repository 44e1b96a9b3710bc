import numpy as np
from tests import datagen, parity
from tests.test_gpu_parity import make_gpu, cfg_for, CASES
from oracle import pyoracle as po
n, p, m, P, family, lik_sigma, kw, tunes = CASES[0]
X, y = datagen.small(n, p, seed=n + m)
cfg = cfg_for(y, m, P, p, seed=n, **kw)
S = P - 1
U = len(tunes) * m
orc = po.OracleSampler(X.astype(np.float64), y, **cfg)
orc.set_likelihood(family, [lik_sigma]); orc.set_rng(po.RNG_SEQ, cfg["seed"])
tape = po.TapeBuffers(U, 64, P); tr = po.TraceBuffers(U, 64, P)
orc.attach_tape(tape); orc.attach_trace(tr)
for t in tunes: orc.step(t)
dev = make_gpu()(X, y, cfg)
dev.set_likelihood_code(family, [lik_sigma]); dev.set_rng_tape(tape.slot, tape.round, tape.upd); dev.trace_enable(U, 64)
for t in tunes: dev.step(t)
ds, dr, du, n_upd, ovf = dev.trace_read()
bad = 0
for u in range(n_upd):
    a, b = tr.upd[u], du[u]
    if not np.allclose(a[4:4+P], b[4:4+P], rtol=1e-6) or a[0] != b[0] or a[1] != b[1]:
        bad += 1
        if bad <= 5:
            print("update", u, "rounds", a[0], b[0], "sel", a[1], b[1], "tree", a[2], b[2])
            print("  ref W", a[4:4+P][:6], " dev W", b[4:4+P][:6], " dev Wn", b[4+P:4+2*P][:4])
            print("  status ref", tr.slot[u,0,:,0], "\n  status dev", ds[u,0,:,0])
print("bad updates", bad, "of", n_upd, "profile", dev.profile()["speculative_root_passes"])
