"""Replay-mode parity harness shared by the host-simulation tests (CPU) and the GPU tests.

Protocol (BASELINE.json north_star): the oracle runs with its sequential
generator and records a typed RNG tape plus a decision trace; the implementation under test replays
the tape.  Bit-exact: status/node/variable/split of every proposal, child counts, column min/max,
ancestors, selected particle, round counts, tree structure, leaf_indices.  Within `rtol` (default
1e-6 relative, the tolerance north_star states): log-weights, leaf values, predictions.
"""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po

RTOL = 1e-6
EPS = float(np.finfo(np.float64).eps)

# Largest discrepancies seen by the last assert_traces_match call, as multiples of the tolerance that was applied
# (tests print / record them so the margin is visible in the GPU logs).
LAST_MARGINS = {}


def weight_tolerance(n: int, wmax: float) -> float:
    """Absolute agreement to expect between the reference's arithmetic and the device's on one log-weight of n rows.

    What differs is summation order.  The reference adds a node's `others` values one by one (the fold of
    smc.rs:201-217) and its log-likelihood terms one by one (weight.rs:26-30 -> the logp loop); a sequential f64 sum of k
    terms is off by up to k * eps / 2 relative, and early in a chain — where the running predictions are piecewise
    constant, so every addend repeats the same rounding — the error IS systematic, not a random walk: measured against
    an extended-precision evaluation the oracle's child sums are off by 0.02-0.06 * n * eps relative (n = 1e3 .. 4e5),
    while the device (fixed-order tree reduction of per-block partial sums) stays at a few ulps.  Leaf values inherit the
    error of the child sums, the running array A inherits it from the leaf values, and the log-likelihood turns a
    relative shift d of the leaf means into an absolute change of order d * |w|.  Hence the bound: 0.25 * n * eps * |w|
    (4x the largest measured coefficient) plus 32 ulps for the closing arithmetic — 8e-4 absolute on log-weights of
    magnitude 1.5e7 at BASELINE C3, i.e. 5e-11 relative, four orders of magnitude inside north_star's 1e-6, which is what
    makes the bit-exact comparison of ancestors and selections meaningful.
    """
    return (0.25 * float(max(n, 1)) + 32.0) * EPS * max(abs(float(wmax)), 1.0)


def _check_logw(got, ref, n, what):
    """Log-weights: north_star's 1e-6 relative AND the absolute bound that follows from the summation error."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    live = np.isfinite(ref)
    np.testing.assert_array_equal(np.isfinite(got), live, err_msg=what + " (finite pattern)")
    if not live.any():
        return 0.0
    wtol = weight_tolerance(n, np.max(np.abs(ref[live])))
    d = np.max(np.abs(got[live] - ref[live]))
    assert d <= wtol, f"{what}: log-weight differs by {d:.3e} > {wtol:.3e} (n={n})"
    np.testing.assert_allclose(got[live], ref[live], rtol=RTOL, atol=0, err_msg=what)
    LAST_MARGINS["logw"] = max(LAST_MARGINS.get("logw", 0.0), d / wtol)
    return wtol


def _check_normalised(got, ref, wtol, P, what):
    """Normalised weights exp(w - max) / sum: an absolute error d in the log-weights is a relative error <= 2 d here
    (numerator and denominator), plus the rounding of exp / the P-term sum / the division."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    rt = 2.0 * wtol + 8.0 * (P + 4) * EPS
    err = np.abs(got - ref)
    lim = rt * np.abs(ref) + 1e-300
    bad = err > lim
    assert not bad.any(), f"{what}: normalised weight off by {np.max(err / lim):.2f}x the bound {rt:.3e}"
    LAST_MARGINS["wnorm"] = max(LAST_MARGINS.get("wnorm", 0.0), float(np.max(err / lim)))


def assert_traces_match(ref, got, n_upd, S, P, rtol=RTOL, n=None):
    """ref/got: (slot, round, upd) arrays in the layout of include/pgbart.h; n = rows behind every log-weight."""
    rs, rr, ru = ref
    gs, gr, gu = got
    LAST_MARGINS.clear()
    if n is None:
        n = 1 << 20          # callers that do not know n get the loosest bound any test here needs
    for u in range(n_upd):
        assert int(ru[u, 0]) == int(gu[u, 0]), f"update {u}: rounds {ru[u, 0]} vs {gu[u, 0]}"
        assert ru[u, 1] == gu[u, 1], f"update {u}: selected particle {ru[u, 1]} vs {gu[u, 1]}"
        assert ru[u, 2] == gu[u, 2] and ru[u, 3] == gu[u, 3], f"update {u}: tree/acceptance"
        wtol = _check_logw(gu[u, 4:4 + P], ru[u, 4:4 + P], n, f"update {u} log W")
        _check_normalised(gu[u, 4 + P:4 + 2 * P], ru[u, 4 + P:4 + 2 * P], wtol, P, f"update {u} normalised W")
        for r in range(int(ru[u, 0])):
            a, b = rs[u, r], gs[u, r]
            for f, name in ((0, "status"), (1, "node"), (2, "var"), (3, "split"), (6, "cntL"), (7, "cntR"),
                            (9, "min"), (10, "max")):
                np.testing.assert_array_equal(b[:, f], a[:, f], err_msg=f"update {u} round {r} {name}")
            for f, name in ((4, "vL"), (5, "vR"), (11, "SoL")):
                np.testing.assert_allclose(b[:, f], a[:, f], rtol=rtol, atol=1e-9, err_msg=f"update {u} round {r} {name}")
            np.testing.assert_array_equal(np.isnan(b[:, 8]), np.isnan(a[:, 8]), err_msg=f"update {u} round {r} logw pattern")
            grown = ~np.isnan(a[:, 8])
            wt = _check_logw(b[grown, 8], a[grown, 8], n, f"update {u} round {r} logw") if grown.any() else weight_tolerance(n, 1.0)
            np.testing.assert_array_equal(gr[u, r, 2 + S:2 + 2 * S], rr[u, r, 2 + S:2 + 2 * S],
                                          err_msg=f"update {u} round {r} ancestors")
            # the weight vector also carries stale entries (normalised weights of earlier rounds, Q1): same bound
            _check_normalised(gr[u, r, 2:2 + S], rr[u, r, 2:2 + S], max(wt, wtol), S, f"update {u} round {r} normalised weights")
            assert gr[u, r, 1] == rr[u, r, 1]


def assert_forest_matches(orc, dev, m, rtol=RTOL, check_leaf_indices=True):
    for t in range(m):
        osv, osp, olv = orc.tree(t)
        dsv, dsp, dlv = dev.tree(t)
        np.testing.assert_array_equal(dsv, osv, err_msg=f"tree {t} split_var")
        np.testing.assert_array_equal(np.nan_to_num(dsp, nan=-7e300), np.nan_to_num(osp, nan=-7e300),
                                      err_msg=f"tree {t} split_val")   # bit-exact incl. NaN placement
        np.testing.assert_array_equal(np.isnan(dlv), np.isnan(olv), err_msg=f"tree {t} leaf NaN pattern")
        np.testing.assert_allclose(np.nan_to_num(dlv), np.nan_to_num(olv), rtol=rtol, atol=1e-12, err_msg=f"tree {t} leaf_val")
        if check_leaf_indices:
            np.testing.assert_array_equal(dev.leaf_indices(t), orc.leaf_indices(t), err_msg=f"tree {t} leaf_indices")


def replay_parity(make_dev, X, y, cfg, family, lik_sigma, tunes, r_cap=24, rtol=RTOL, expect_error=None):
    """Run oracle (recording) and implementation (replaying) side by side.

    make_dev(X, y, cfg) -> object with set_likelihood_code / set_rng_tape / trace_enable / trace_read / step /
    tree / leaf_indices / info.  cfg: dict of PyBartSettings-like fields.  Returns the oracle trace.
    """
    m, P = cfg["n_trees"], cfg["n_particles"]
    S = P - 1
    U = len(tunes) * m
    orc = po.OracleSampler(X.astype(np.float64), y, **cfg)
    orc.set_likelihood(family, [lik_sigma])
    orc.set_rng(po.RNG_SEQ, cfg.get("seed", 0))
    tape = po.TapeBuffers(U, r_cap, P)
    tr = po.TraceBuffers(U, r_cap, P)
    orc.attach_tape(tape)
    orc.attach_trace(tr)
    opreds = []
    oerr = None
    try:
        for tune in tunes:
            opreds.append(orc.step(tune))
    except RuntimeError as e:
        oerr = str(e)
    assert orc.overflowed() == 0, "oracle trace/tape overflow: raise r_cap"

    dev = make_dev(X, y, cfg)
    dev.set_likelihood_code(family, [lik_sigma])
    dev.set_rng_tape(tape.slot, tape.round, tape.upd)
    dev.trace_enable(U, r_cap)
    derr = None
    dpreds = []
    try:
        for tune in tunes:
            dpreds.append(dev.step(tune))
    except RuntimeError as e:
        derr = str(e)
    if expect_error is not None:
        assert oerr is not None and expect_error in oerr, f"oracle error: {oerr}"
        assert derr is not None and expect_error in derr, f"device error: {derr}"
        return tr
    assert oerr is None and derr is None, (oerr, derr)
    ds, dr, du, n_upd, ovf = dev.trace_read()
    assert not ovf
    assert n_upd == orc.trace_count()
    assert_traces_match((tr.slot, tr.round, tr.upd), (ds, dr, du), n_upd, S, P, rtol, n=X.shape[0])
    for a, b in zip(opreds, dpreds):
        np.testing.assert_allclose(b, a, rtol=rtol, atol=0, err_msg="predictions")
    assert_forest_matches(orc, dev, m, rtol)
    oll, oacc, odep = orc.info()
    dll, dacc, ddep = dev.info()
    assert oacc == dacc
    np.testing.assert_array_equal(ddep, odep)
    np.testing.assert_allclose(dll, oll, rtol=1e-5, atol=1e-12)
    return tr


def philox_parity(make_dev, X, y, cfg, family, lik_sigma, tunes, r_cap=24, rtol=RTOL):
    """Free-running Philox: oracle and implementation generate the same draws independently."""
    m, P = cfg["n_trees"], cfg["n_particles"]
    S = P - 1
    U = len(tunes) * m
    seed = cfg.get("seed", 0)
    orc = po.OracleSampler(X.astype(np.float64), y, **cfg)
    orc.set_likelihood(family, [lik_sigma])
    orc.set_rng(po.RNG_PHILOX, seed)
    tr = po.TraceBuffers(U, r_cap, P)
    orc.attach_trace(tr)
    dev = make_dev(X, y, cfg)
    dev.set_likelihood_code(family, [lik_sigma])
    dev.set_rng_philox(seed)
    dev.trace_enable(U, r_cap)
    for tune in tunes:
        a = orc.step(tune)
        b = dev.step(tune)
        np.testing.assert_allclose(b, a, rtol=rtol, atol=0)
    ds, dr, du, n_upd, ovf = dev.trace_read()
    assert not ovf and orc.overflowed() == 0
    assert_traces_match((tr.slot, tr.round, tr.upd), (ds, dr, du), n_upd, S, P, rtol, n=X.shape[0])
    assert_forest_matches(orc, dev, m, rtol)
    return tr


# ---------------------------------------------------------------------------------------------------------
# Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): replay without the oracle.

GOLDEN_NAMES = ["gauss_P20", "gauss_deep_P3", "bern_P6", "gauss_noprior_batch"]


def load_golden(name):
    import ast
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    z = np.load(path)
    g = {k: z[k] for k in z.files if k != "meta"}
    g["meta"] = ast.literal_eval(str(z["meta"]))
    return g


class GoldenForest:
    """tree()/leaf_indices() view of a fixture, for assert_forest_matches."""

    def __init__(self, g):
        self.g = g

    def tree(self, t):
        sv = self.g["split_var"][t]
        # stored padded to the largest tree of the fixture: cut back to this tree's own length
        live = np.nonzero((sv >= 0) | ~np.isnan(self.g["leaf_val"][t]))[0]
        size = int(live.max()) + 1 if live.size else 1
        return sv[:size].astype(np.int32), self.g["split_val"][t][:size], self.g["leaf_val"][t][:size]

    def leaf_indices(self, t):
        return self.g["leaf_indices"][t].astype(np.uint32)


def golden_parity(make_dev, g, rtol=RTOL):
    """Replay a fixture's tape through `make_dev`'s implementation and compare with the stored trace."""
    meta = g["meta"]
    cfg = {k: meta[k] for k in ("n_trees", "n_particles", "max_depth", "alpha", "beta", "sigma", "split_prior",
                                "batch_tune", "batch_post", "seed")}
    m, P = cfg["n_trees"], cfg["n_particles"]
    S = P - 1
    dev = make_dev(g["X"], g["y"], cfg)
    dev.set_likelihood_code(meta["family"], [meta["lik_sigma"]])
    dev.set_rng_tape(g["tape_slot"], g["tape_round"], g["tape_upd"])
    U = g["tape_upd"].shape[0]
    dev.trace_enable(U, meta["r_cap"])
    preds = np.stack([dev.step(t) for t in meta["tunes"]])
    ds, dr, du, n_upd, ovf = dev.trace_read()
    assert not ovf and n_upd == int(g["n_upd"])
    assert_traces_match((g["trace_slot"], g["trace_round"], g["trace_upd"]), (ds, dr, du), n_upd, S, P, rtol, n=g["X"].shape[0])
    np.testing.assert_allclose(preds, g["preds"], rtol=rtol, atol=0, err_msg="predictions")
    gf = GoldenForest(g)
    for t in range(m):
        osv, osp, olv = gf.tree(t)
        dsv, dsp, dlv = dev.tree(t)
        k = min(len(osv), len(dsv))       # implementations may trim trailing unused nodes differently
        assert np.all(dsv[k:] < 0) and np.all(osv[k:] < 0)
        np.testing.assert_array_equal(dsv[:k], osv[:k], err_msg=f"tree {t} split_var")
        np.testing.assert_array_equal(np.nan_to_num(dsp[:k], nan=-7e300), np.nan_to_num(osp[:k], nan=-7e300))
        np.testing.assert_array_equal(np.isnan(dlv[:k]), np.isnan(olv[:k]))
        np.testing.assert_allclose(np.nan_to_num(dlv[:k]), np.nan_to_num(olv[:k]), rtol=rtol, atol=1e-12)
        np.testing.assert_array_equal(dev.leaf_indices(t), gf.leaf_indices(t), err_msg=f"tree {t} leaf_indices")
    dll, dacc, ddep = dev.info()
    assert dacc == int(g["info_acc"])
    np.testing.assert_array_equal(ddep, g["info_depths"])
