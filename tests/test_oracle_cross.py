"""Cross-check the C++ oracle against an independent pure-Python restatement on tiny cases.

The oracle runs with its sequential generator and records a typed RNG tape;
the Python restatement replays that tape.  Decisions must be identical, values equal to 1e-12.
Also checks that the oracle replaying its own tape reproduces itself, and the reference's
stale-weight consequences.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import datagen
from tests.py_restatement import PyRestatement


def compare_traces(a, b, n_upd, S, P, rtol=1e-12, atol=1e-12):
    """a, b: TraceBuffers.  Bit-exact decisions, close values."""
    for u in range(n_upd):
        ra, rb = int(a.upd[u, 0]), int(b.upd[u, 0])
        assert ra == rb, f"update {u}: rounds {ra} vs {rb}"
        assert a.upd[u, 1] == b.upd[u, 1], f"update {u}: selected differs"
        assert a.upd[u, 2] == b.upd[u, 2] and a.upd[u, 3] == b.upd[u, 3]
        np.testing.assert_allclose(a.upd[u, 4:4 + P], b.upd[u, 4:4 + P], rtol=rtol, atol=atol)
        np.testing.assert_allclose(a.upd[u, 4 + P:], b.upd[u, 4 + P:], rtol=1e-9, atol=1e-300)
        for r in range(ra):
            sa, sb = a.slot[u, r], b.slot[u, r]
            # exact: status, node, var, split, counts, lo, hi
            for f in (0, 1, 2, 3, 6, 7, 9, 10):
                np.testing.assert_array_equal(sa[:, f], sb[:, f], err_msg=f"update {u} round {r} field {f}")
            for f in (4, 5, 8, 11):
                np.testing.assert_allclose(sa[:, f], sb[:, f], rtol=rtol, atol=atol,
                                           err_msg=f"update {u} round {r} field {f}")
            np.testing.assert_array_equal(a.round[u, r, 2 + S:], b.round[u, r, 2 + S:],
                                          err_msg=f"update {u} round {r} ancestors")
            np.testing.assert_allclose(a.round[u, r, 2:2 + S], b.round[u, r, 2:2 + S], rtol=1e-9, atol=1e-300)
            assert a.round[u, r, 1] == b.round[u, r, 1]


CASES = [
    # n, p, m, P, family, lik_sigma, use_prior, seed, steps
    (40, 3, 4, 5, 0, 1.0, True, 0, 3),
    (64, 5, 6, 8, 0, 0.5, True, 1, 2),
    (30, 2, 3, 2, 0, 1.0, False, 2, 4),     # P=2: one growing particle, trees grow deep
    (50, 4, 5, 6, 1, 1.0, True, 3, 2),      # Bernoulli-logit
    (33, 3, 4, 4, 2, 1.0, True, 4, 2),      # bench mock likelihood
    (25, 2, 3, 3, 0, 0.01, True, 5, 3),     # tiny sigma: positive log-likelihoods beat stale weights
]


@pytest.mark.parametrize("n,p,m,P,family,lik_sigma,use_prior,seed,steps", CASES)
def test_oracle_matches_python_restatement(n, p, m, P, family, lik_sigma, use_prior, seed, steps):
    X, y = datagen.small(n, p, seed=seed, bernoulli=(family == 1))
    prm = datagen.pgbart_params(y, m, p)
    prior = prm["split_prior"] if use_prior else []
    kw = dict(n_trees=m, n_particles=P, max_depth=4, alpha=0.95, beta=1.0, sigma=prm["sigma"],
              split_prior=prior, batch_tune=1.0, batch_post=0.5)
    U, R = steps * m, 40
    orc = po.OracleSampler(X.astype(np.float64), y, seed=seed, **kw)
    orc.set_likelihood(family, [lik_sigma])
    orc.set_rng(po.RNG_SEQ, seed)
    tape = po.TapeBuffers(U, R, P)
    tr_o = po.TraceBuffers(U, R, P)
    orc.attach_tape(tape)
    orc.attach_trace(tr_o)

    py = PyRestatement(X, y, family=family, lik_sigma=lik_sigma, **kw)
    tr_p = po.TraceBuffers(U, R, P)
    try:
        for s in range(steps):
            tune = s == 0
            po_pred = orc.step(tune)
            py_pred = py.step(tape, tr_p, tune)
            np.testing.assert_allclose(po_pred, py_pred, rtol=1e-12, atol=1e-12)
    except RuntimeError as e:  # both must hit the reference's depth panic at the same place
        assert "exceed maximum capacity" in str(e)
        return
    assert orc.overflowed() == 0
    n_upd = orc.trace_count()
    assert n_upd == py.upd
    compare_traces(tr_o, tr_p, n_upd, P - 1, P)


def test_oracle_replays_its_own_tape():
    n, p, m, P = 80, 5, 6, 10
    X, y = datagen.friedman(n, p, seed=7)
    prm = datagen.pgbart_params(y, m, p)
    kw = dict(n_trees=m, n_particles=P, max_depth=prm["max_depth"], alpha=0.95, beta=2.0, sigma=prm["sigma"],
              split_prior=prm["split_prior"], batch_tune=1.0, batch_post=1.0)
    U, R = 3 * m, 16
    a = po.OracleSampler(X.astype(np.float64), y, seed=3, **kw)
    a.set_rng(po.RNG_SEQ, 3)
    tape, tra = po.TapeBuffers(U, R, P), po.TraceBuffers(U, R, P)
    a.attach_tape(tape)
    a.attach_trace(tra)
    pa = [a.step() for _ in range(3)]
    b = po.OracleSampler(X.astype(np.float64), y, seed=99, **kw)
    b.set_rng(po.RNG_REPLAY)
    trb = po.TraceBuffers(U, R, P)
    b.attach_tape(tape)
    b.attach_trace(trb)
    pb = [b.step() for _ in range(3)]
    for u, v in zip(pa, pb):
        np.testing.assert_array_equal(u, v)
    np.testing.assert_array_equal(np.nan_to_num(tra.slot, nan=-777.0), np.nan_to_num(trb.slot, nan=-777.0))
    np.testing.assert_array_equal(tra.upd, trb.upd)


def test_stale_weight_rule_consequences():
    """with alpha=.95, beta=2 a tree update returns a stump unless every slot grew in round 1;
    trees never exceed one split; the round count is 1 or 3 (rarely more)."""
    n, p, m, P = 200, 5, 50, 20
    X, y = datagen.friedman(n, p, seed=11)
    prm = datagen.pgbart_params(y, m, p)
    o = po.OracleSampler(X.astype(np.float64), y, n_trees=m, n_particles=P, max_depth=prm["max_depth"],
                         alpha=0.95, beta=2.0, sigma=prm["sigma"], split_prior=prm["split_prior"],
                         batch_tune=1.0, batch_post=1.0, seed=0)
    o.set_rng(po.RNG_PHILOX, 0)
    U = 6 * m
    tr = po.TraceBuffers(U, 12, P)
    o.attach_trace(tr)
    for _ in range(6):
        o.step()
    assert o.overflowed() == 0
    rounds = tr.upd[:U, 0]
    status1 = tr.slot[:U, 0, :, 0]
    all_grew = np.all(status1 == 4, axis=1)
    for u in range(U):
        t = int(tr.upd[u, 2])
        if not all_grew[u]:
            assert rounds[u] == 1 and tr.upd[u, 1] >= 0
    sizes = [o.tree(t)[0].size for t in range(m)]
    assert max(sizes) <= 3
    frac_stump_updates = np.mean(~all_grew)
    assert abs(frac_stump_updates - (1 - 0.95 ** (P - 1))) < 0.1
