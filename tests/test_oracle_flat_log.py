"""The oracle's FLAT random source: a sequential log of what a reference run consumed (raw generator words for uniform /
range draws, the value of every standard normal) — the format the Rust-side dumper of INTEGRATION.md section 1b writes.
No Rust toolchain exists here, so the log is produced by the oracle's own sequential generator and must reproduce that
run exactly, typed tape included (which is what the device replays)."""
import numpy as np

from oracle import pyoracle as po
from tests import datagen
from tests.test_sim_parity import cfg_for


def _run(X, y, cfg, mode, flat=None, rec=None, steps=3):
    m, P = cfg["n_trees"], cfg["n_particles"]
    o = po.OracleSampler(X.astype(np.float64), y, **cfg)
    o.set_likelihood(po.FAM_GAUSSIAN, [1.0])
    o.set_rng(mode, cfg["seed"])
    tape, tr = po.TapeBuffers(steps * m, 32, P), po.TraceBuffers(steps * m, 32, P)
    o.attach_tape(tape)
    o.attach_trace(tr)
    if flat is not None:
        o.attach_flat(flat)
    if rec is not None:
        o.record_flat(rec)
    preds = [o.step(True) for _ in range(steps)]
    return o, tape, tr, preds


def test_flat_log_reproduces_the_sequential_run():
    X, y = datagen.small(800, 4, seed=3)
    cfg = cfg_for(y, 6, 5, 4, seed=9, beta=0.8, max_depth=10)
    log = np.zeros(100_000)
    o1, tape1, tr1, p1 = _run(X, y, cfg, po.RNG_SEQ, rec=log)
    n = o1.flat_count()
    assert 0 < n < log.size
    o2, tape2, tr2, p2 = _run(X, y, cfg, po.RNG_FLAT, flat=log[:n])
    assert o2.flat_count() == n and o2.overflowed() == 0           # every logged word consumed, none missing
    for a, b in zip(p1, p2):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(tape1.slot, tape2.slot)           # the typed tape the device replays
    np.testing.assert_array_equal(tape1.round, tape2.round)
    np.testing.assert_array_equal(tape1.upd, tape2.upd)
    np.testing.assert_array_equal(np.nan_to_num(tr1.slot, nan=-7e300), np.nan_to_num(tr2.slot, nan=-7e300))
    # a truncated log is an error, not a silent divergence
    o3 = po.OracleSampler(X.astype(np.float64), y, **cfg)
    o3.set_likelihood(po.FAM_GAUSSIAN, [1.0])
    o3.set_rng(po.RNG_FLAT, cfg["seed"])
    o3.attach_flat(log[: n // 2])
    failed = False
    try:
        for _ in range(3):
            o3.step(True)
    except RuntimeError:
        failed = True
    assert failed
