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


@pytest.mark.parametrize("case", range(10))
def test_sim_replay_parity_randomized(case):
    """Seeded random configurations (sizes, particle counts on both sides of 32, families, priors, batch fractions, weight
    modes, grid sizes): the same generator feeds the GPU suite with larger n (tests/test_gpu_parity.py)."""
    n, p, m, P, family, lik_sigma, kw, tunes, pg_correct, grid = datagen.random_cases(10, 2024, 20, 300)[case]
    X, y = datagen.small(n, p, seed=1000 + case, bernoulli=(family == 1))
    cfg = cfg_for(y, m, P, p, seed=case, **kw)
    if pg_correct:
        cfg["pg_correct"] = True
    parity.replay_parity(make_sim(grid), X, y, cfg, family, lik_sigma, tunes, r_cap=128)


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


# ---- modes beyond the reference's live behaviour, each against its own oracle mode -------------------

def _categorical(n, p, seed, cats=(0, 1, 2)):
    """Columns 0 .. len(cats)-1 hold integer codes (column j: cats[j] + 2 distinct values, shifted by j - 1), the rest U[0, 1)."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, p), dtype=np.float32)
    for j, extra in enumerate(cats):
        X[:, j] = (rng.integers(0, 2 + extra, n) + (j - 1)).astype(np.float32)
    y = (X[:, 0] * 1.5 - (X[:, 1] > 0) + np.sin(3 * X[:, -1]) + 0.3 * rng.standard_normal(n)).astype(np.float32).astype(np.float64)
    return X, y


PG_CORRECT_CASES = [
    (200, 4, 6, 8, 0, 1.0, dict(), [True, True]),
    (150, 3, 5, 4, 1, 1.0, dict(beta=1.0, max_depth=8), [True, True]),
    (120, 3, 4, 40, 0, 0.5, dict(), [True]),                    # P > 32
    (90, 3, 4, 3, 2, 1.0, dict(beta=0.7, max_depth=9), [True, True, True]),
]


@pytest.mark.parametrize("n,p,m,P,family,lik_sigma,kw,tunes", PG_CORRECT_CASES)
def test_sim_pg_correct_mode(n, p, m, P, family, lik_sigma, kw, tunes):
    """weight_mode pg_correct: log-weights carried with the particles, the tree being replaced as reference particle.
    Unlike the reference mode (Q1) trees grow past one split and the old tree can be kept."""
    X, y = datagen.small(n, p, seed=n + P, bernoulli=(family == 1))
    cfg = cfg_for(y, m, P, p, seed=n, **kw)
    cfg["pg_correct"] = True
    tr = parity.replay_parity(make_sim(2), X, y, cfg, family, lik_sigma, tunes, r_cap=96)
    U = len(tunes) * m
    assert np.any(tr.upd[:U, 0] > 3), "no tree update ran more than three rounds: the mode did not take effect"


@pytest.mark.parametrize("P,pg_correct", [(6, False), (5, True), (40, True)])
def test_sim_onehot_columns(P, pg_correct):
    """OneHotSplit columns (splitting.rs:72-104): the split value is one of the node's distinct integer codes; the
    partition still tests x < code (Q7).  With pg_correct trees get deep enough for non-root nodes to draw codes."""
    n, p, m = 240, 5, 5
    X, y = _categorical(n, p, seed=P)
    cfg = cfg_for(y, m, P, p, seed=P, beta=0.8, max_depth=9)
    cfg["onehot"] = [1, 1, 1, 0, 0]
    if pg_correct:
        cfg["pg_correct"] = True
    tr = parity.replay_parity(make_sim(3), X, y, cfg, 0, 1.0, [True, True], r_cap=96)
    st = tr.slot[:2 * m]
    onehot_accepts = (st[..., 0] == 4) & (st[..., 2] <= 2)
    assert onehot_accepts.any()
    assert np.all(st[..., 3][onehot_accepts] == np.round(st[..., 3][onehot_accepts]))      # split values are integer codes
    if pg_correct:
        assert (onehot_accepts & (st[..., 1] > 0)).any(), "no OneHot split below the root: the presence-mask pass was not exercised"
