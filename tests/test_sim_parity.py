"""Host simulation of the device sweep (product serial logic + loop-based passes) vs the oracle.

Runs in the GPU-less container: checks the state machine, stale-weight rule, resampling, slot-table
copies, job grouping, partition offsets, Bernoulli pass ordering, error paths.  The GPU tests
(tests/test_gpu_parity.py) run the same harness through the C ABI on the real kernels.
"""
import numpy as np
import pytest

from tests import datagen, parity
from tests.sim_lib import SimSampler, lib


def make_sim(grid):
    def mk(X, y, cfg):
        return SimSampler(X, y, grid=grid, **cfg)
    return mk


def cfg_for(y, m, P, p, *, max_depth=None, alpha=0.95, beta=2.0, prior=True, batch=(1.0, 1.0), seed=0):
    prm = datagen.pgbart_params(y, m, p, alpha, beta)
    return dict(n_trees=m, n_particles=P, max_depth=prm["max_depth"] if max_depth is None else max_depth, alpha=alpha,
                beta=beta, sigma=prm["sigma"], split_prior=prm["split_prior"] if prior else [], batch_tune=batch[0],
                batch_post=batch[1], seed=seed)


CASES = [
    # n, p, m, P, family, lik_sigma, kwargs, tunes, grid
    (200, 5, 10, 20, 0, 1.0, dict(), [True, False], 3),
    (257, 6, 8, 8, 0, 0.7, dict(batch=(0.5, 0.25)), [True, True, False, False], 2),
    (64, 3, 5, 2, 0, 1.0, dict(beta=1.0, max_depth=10, prior=False), [True, True, True], 1),   # P=2: deep trees
    (90, 4, 6, 3, 0, 1.0, dict(beta=0.5, max_depth=10), [True, True], 4),                        # deeper trees, few slots
    (120, 5, 6, 6, 1, 1.0, dict(), [True, False], 2),                                           # Bernoulli-logit
    (75, 3, 4, 2, 1, 1.0, dict(beta=1.0, max_depth=10), [True, True], 3),                        # Bernoulli, deep
    (100, 5, 7, 5, 2, 1.0, dict(), [True, True], 2),                                            # bench mock likelihood
    (40, 2, 3, 4, 0, 0.01, dict(), [True, True, True], 2),                                      # positive log-liks (tiny sigma)
    (33, 2, 4, 6, 0, 1.0, dict(), [True], 5),                                                   # more warps than rows/4
    (1, 2, 3, 4, 0, 1.0, dict(), [True], 1),                                                    # single row: min>=max rejects
    (150, 4, 3, 40, 1, 1.0, dict(alpha=0.9999, beta=1.5, max_depth=6), [True], 2),              # P > 32, Bernoulli, every slot grows
    (130, 3, 3, 70, 0, 1.0, dict(), [True, False], 3),                                          # P > 32, default prior (stale zero weights)
]


@pytest.mark.parametrize("n,p,m,P,family,lik_sigma,kw,tunes,grid", CASES)
def test_sim_replay_parity(n, p, m, P, family, lik_sigma, kw, tunes, grid):
    X, y = datagen.small(n, p, seed=n + m, bernoulli=(family == 1))
    cfg = cfg_for(y, m, P, p, seed=n, **kw)
    parity.replay_parity(make_sim(grid), X, y, cfg, family, lik_sigma, tunes, r_cap=64)


def test_sim_philox_parity():
    X, y = datagen.friedman(300, 5, seed=2)
    cfg = cfg_for(y, 12, 10, 5, seed=5)
    parity.philox_parity(make_sim(3), X, y, cfg, 0, 1.0, [True, False, False])


def test_sim_constant_column_and_duplicates():
    # a constant column (min>=max reject) and a column with many ties
    rng = np.random.default_rng(3)
    X = rng.random((150, 4), dtype=np.float32)
    X[:, 3] = 0.25
    X[:, 2] = np.round(X[:, 2] * 3) / 3
    y = (X[:, 0] * 3 + rng.standard_normal(150) * 0.3).astype(np.float32).astype(np.float64)
    cfg = cfg_for(y, 8, 6, 4, seed=1)
    tr = parity.replay_parity(make_sim(2), X, y, cfg, 0, 1.0, [True, True, True], r_cap=64)
    assert np.any(tr.slot[..., 0] == 3), "the constant column should have produced min>=max rejections"


def test_sim_depth_overflow_matches_reference_panic():
    # max_depth=1 with one growing particle and alpha close to 1: a depth-1 node gets split -> reference panics
    X, y = datagen.small(60, 2, seed=9)
    cfg = cfg_for(y, 3, 2, 2, max_depth=1, alpha=0.999, beta=0.01, seed=4)
    parity.replay_parity(make_sim(2), X, y, cfg, 0, 1.0, [True, True, True, True], r_cap=64,
                         expect_error="exceed maximum capacity")


def test_helpers_thr32_keys_ranges():
    L = lib()
    rng = np.random.default_rng(0)
    # (double)x < s  <=>  x < thr32(s)
    for s in list(rng.standard_normal(200)) + [0.0, -0.0, 1e-50, -1e-50, 1e30, -1e30, 0.1, float(np.float32(0.1))]:
        t = np.float32(L.pgsim_thr32(float(s)))
        xs = np.array([np.float32(s), np.nextafter(np.float32(s), np.float32(np.inf)),
                       np.nextafter(np.float32(s), np.float32(-np.inf)), t], dtype=np.float32)
        for x in xs:
            assert (float(x) < s) == bool(x < t), (s, x, t)
    vals = np.sort(rng.standard_normal(100).astype(np.float32))
    keys = [L.pgsim_f2key(float(v)) for v in vals]
    assert keys == sorted(keys)
    for v in list(vals) + [np.float32(0.0), np.float32(-0.0)]:
        assert np.float32(L.pgsim_key2f(L.pgsim_f2key(float(v)))) == v
    import ctypes as C
    for Lh, W, g in [(1000, 24, 4), (33, 40, 4), (7, 16, 1), (0, 8, 1), (100000, 2368, 4)]:
        cover = 0
        prev = 0
        for gw in range(W):
            b, e = C.c_uint64(), C.c_uint64()
            L.pgsim_warp_range(Lh, gw, W, g, C.byref(b), C.byref(e))
            assert b.value == prev and e.value >= b.value and (b.value % g == 0 or b.value == Lh)
            prev = e.value
            cover += e.value - b.value
        assert cover == Lh and prev == Lh
