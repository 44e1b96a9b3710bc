"""Generate the golden replay fixtures in this directory.

    python tests/golden/make_golden.py

Each fixture is one seeded run of the CPU oracle (oracle/pgbart_oracle.cpp, sequential generator): the inputs
(X, y stored verbatim so the fixture does not depend on numpy's generators), the settings, the typed RNG tape it
drew, its decision trace (layout of include/pgbart.h), the predictions after every step,
and the final forest.  The Rust reference cannot be run in this image (no cargo/rustc), so these vectors come
from the restatement, which tests/test_oracle_kat.py pins to the reference's own unit-test vectors and
tests/test_oracle_cross.py pins to an independent Python restatement.  The GPU tests replay the tape through
the C ABI and compare with the stored trace without running the oracle at all; the CPU tests re-run the oracle
and require the stored bytes back, so a change in the oracle's arithmetic cannot go unnoticed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po      # noqa: E402
from tests import datagen              # noqa: E402

R_CAP = 64

# name, n, p, m, P, family, lik_sigma, bernoulli data, settings overrides, tunes
FIXTURES = [
    ("gauss_P20", 600, 6, 6, 20, po.FAM_GAUSSIAN, 1.0, False, dict(), [True, False]),
    ("gauss_deep_P3", 500, 3, 5, 3, po.FAM_GAUSSIAN, 1.0, False, dict(beta=0.5, max_depth=10), [True, True]),
    ("bern_P6", 800, 4, 5, 6, po.FAM_BERNOULLI_LOGIT, 1.0, True, dict(), [True, False]),
    ("gauss_noprior_batch", 400, 5, 8, 5, po.FAM_GAUSSIAN, 0.5, False, dict(prior=False, batch=(0.5, 0.25)),
     [True, True, False, False]),
]


def settings_for(y, m, P, p, *, alpha=0.95, beta=2.0, max_depth=None, prior=True, batch=(1.0, 1.0), seed=0):
    prm = datagen.pgbart_params(y, m, p, alpha, beta)
    return dict(n_trees=m, n_particles=P, max_depth=prm["max_depth"] if max_depth is None else max_depth, alpha=alpha,
                beta=beta, sigma=float(prm["sigma"]), split_prior=[float(v) for v in prm["split_prior"]] if prior else [],
                batch_tune=batch[0], batch_post=batch[1], seed=seed)


def run_oracle(X, y, cfg, family, lik_sigma, tunes, tape=None):
    """Run the oracle; with `tape` None it draws sequentially and records, else it replays `tape`."""
    m, P = cfg["n_trees"], cfg["n_particles"]
    U = len(tunes) * m
    o = po.OracleSampler(X.astype(np.float64), y, **cfg)
    o.set_likelihood(family, [lik_sigma])
    rec = po.TapeBuffers(U, R_CAP, P)
    if tape is None:
        o.set_rng(po.RNG_SEQ, cfg["seed"])
    else:
        rec.slot[...], rec.round[...], rec.upd[...] = tape
        o.set_rng(po.RNG_REPLAY, cfg["seed"])
    o.attach_tape(rec)
    tr = po.TraceBuffers(U, R_CAP, P)
    o.attach_trace(tr)
    preds = np.stack([np.array(o.step(t), dtype=np.float64) for t in tunes])
    assert o.overflowed() == 0
    n_upd = o.trace_count()
    trees = [o.tree(t) for t in range(m)]
    maxn = max(len(t[0]) for t in trees)
    sv = np.full((m, maxn), -1, dtype=np.int64)
    sp = np.full((m, maxn), np.nan)
    lv = np.full((m, maxn), np.nan)
    for t, (a, b, c) in enumerate(trees):
        sv[t, :len(a)], sp[t, :len(b)], lv[t, :len(c)] = a, b, c
    li = np.stack([np.asarray(o.leaf_indices(t), dtype=np.int64) for t in range(m)])
    ll, acc, dep = o.info()
    o.close()
    return dict(tape_slot=rec.slot, tape_round=rec.round, tape_upd=rec.upd, trace_slot=tr.slot, trace_round=tr.round,
                trace_upd=tr.upd, n_upd=np.int64(n_upd), preds=preds, split_var=sv, split_val=sp, leaf_val=lv,
                leaf_indices=li, info_acc=np.float64(acc), info_depths=np.asarray(dep), info_ll=np.float64(ll))


def main():
    for name, n, p, m, P, family, lik_sigma, bern, kw, tunes in FIXTURES:
        X, y = datagen.small(n, p, seed=len(name) * 7 + n, bernoulli=bern)
        cfg = settings_for(y, m, P, p, seed=n + 1, **kw)
        out = run_oracle(X, y, cfg, family, lik_sigma, tunes)
        meta = dict(cfg)
        meta.update(family=int(family), lik_sigma=float(lik_sigma), tunes=[bool(t) for t in tunes], r_cap=R_CAP)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, X=X, y=y, meta=np.array(repr(meta)), **out)
        print(f"{name}: {int(out['n_upd'])} tree updates -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
