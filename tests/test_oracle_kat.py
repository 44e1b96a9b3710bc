"""Known-answer tests pinning the CPU oracle to the reference's own unit tests.

Each test names the Rust test it ports (paths relative to /root/reference).
These are the only results the reference pins for this path.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _kat_tree(init, n, max_depth, splits):
    L = po.lib()
    k = len(splits)
    leaf = np.array([s[0] for s in splits], dtype=np.uint64)
    var = np.array([s[1] for s in splits], dtype=np.uint32)
    val = np.array([s[2] for s in splits], dtype=np.float64)
    lv = np.array([s[3] for s in splits], dtype=np.float64)
    rv = np.array([s[4] for s in splits], dtype=np.float64)
    cap = 64
    sv = np.full(cap, -7, dtype=np.int32)
    leafval = np.full(cap, np.nan)
    li = np.zeros(n, dtype=np.uint32)
    pred = np.zeros(n)
    size = L.orc_kat_tree(init, n, max_depth, k, _p(leaf), _p(var), _p(val), _p(lv), _p(rv),
                          _p(sv), _p(leafval), _p(li), _p(pred), cap)
    return size, sv, leafval, li, pred


def test_new_tree_is_single_leaf():  # src/tree.rs:205-212
    size, sv, lv, li, _ = _kat_tree(1.5, 10, 5, [])
    assert size == 1 and sv[0] == -1 and lv[0] == 1.5
    assert li.shape == (10,) and np.all(li == 0)


def test_split_node_creates_children():  # src/tree.rs:215-225
    size, sv, lv, _, _ = _kat_tree(0.0, 5, 5, [(0, 0, 2.5, -1.0, 1.0)])
    assert size == 3
    assert sv[0] == 0 and sv[1] == -1 and sv[2] == -1
    assert lv[1] == -1.0 and lv[2] == 1.0 and np.isnan(lv[0])


def test_get_depth():  # src/tree.rs:228-235
    L = po.lib()
    assert [L.orc_kat_get_depth(i) for i in (0, 1, 2, 3, 6)] == [0, 1, 1, 2, 2]


def test_max_nodes_for_depth():  # src/tree.rs:238-245
    L = po.lib()
    assert [L.orc_kat_max_nodes_for_depth(d) for d in (0, 1, 2, 5, 6, 9)] == [1, 3, 7, 63, 127, 1023]


def test_predict_training_root_only():  # src/tree.rs:248-253
    _, _, _, _, pred = _kat_tree(3.14, 4, 5, [])
    assert pred.shape == (4,) and np.all(np.abs(pred - 3.14) < 1e-10)


def test_split_beyond_max_depth_panics():  # src/tree.rs:269-275
    size, *_ = _kat_tree(0.0, 1, 1, [(0, 0, 0.5, -1.0, 1.0), (1, 0, 0.3, -2.0, 2.0)])
    assert size == -1  # the oracle reports the reference's panic as an error


def _apply(x, node, var, split, lv, rv, share):
    L = po.lib()
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    n, p = x.shape
    data = np.zeros(n, dtype=np.uint32)
    sl = np.zeros(4, dtype=np.uint32)
    li = np.zeros(n, dtype=np.uint32)
    uc, stump = C.c_int(), C.c_int()
    rc = L.orc_kat_apply_mutation(_p(x), n, p, 3, node, var, split, lv, rv, int(share), _p(data), _p(sl), _p(li),
                                  C.byref(uc), C.byref(stump))
    assert rc == 0
    return data, sl, li, uc.value, stump.value


def test_leaf_samples_flat_new_and_all_in_data():  # src/particle.rs:157-160, 231-237
    # a fresh particle's CSR data is 0..n (observed through an apply at the root that sends all rows left)
    x = np.arange(8, dtype=np.float64).reshape(8, 1)
    data, sl, _, _, _ = _apply(x, 0, 0, 100.0, 0.0, 0.0, False)
    assert list(data) == list(range(8)) and data.sum() == 8 * 7 // 2
    assert tuple(sl) == (0, 8, 8, 0)


def test_apply_mutation_partitions_correctly():  # src/particle.rs:170-201
    x = np.arange(5, dtype=np.float64).reshape(5, 1)
    data, sl, li, use_count, _ = _apply(x, 0, 0, 2.5, -1.0, 1.0, False)
    left = sorted(data[sl[0]: sl[0] + sl[1]])
    right = sorted(data[sl[2]: sl[2] + sl[3]])
    assert left == [0, 1, 2] and right == [3, 4]
    assert list(li) == [1, 1, 1, 2, 2]
    assert use_count == 1


def test_apply_mutation_cow_unshares():  # src/particle.rs:163-167, 204-228
    x = np.arange(4, dtype=np.float64).reshape(4, 1)
    _, _, li, use_count, original_is_stump = _apply(x, 0, 0, 2.0, -1.0, 1.0, True)
    assert original_is_stump == 1 and use_count == 1
    assert list(li) == [1, 1, 2, 2]


def test_get_leaf_samples_equivalent():  # src/tree.rs:256-267 (leaf_indices -> per-leaf sample sets)
    x = np.array([[0.1], [0.2], [0.7], [0.9]])
    _, _, li, _, _ = _apply(x, 0, 0, 0.5, -1.0, 1.0, False)
    assert list(np.nonzero(li == 1)[0]) == [0, 1] and list(np.nonzero(li == 2)[0]) == [2, 3]


# ---- third-party arithmetic restated in the oracle: published test vectors ----

def test_xoshiro256pp_reference_vector():
    # xoshiro256++ reference implementation, state {1,2,3,4}
    out = np.zeros(10, dtype=np.uint64)
    po.lib().orc_kat_xoshiro(1, 2, 3, 4, 10, _p(out))
    assert list(out) == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                         9973669472204895162, 14011001112246962877, 12406186145184390807,
                         15849039046786891736, 10450023813501588000]


@pytest.mark.parametrize("ctr,key,exp", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox4x32_10_random123_vectors(ctr, key, exp):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    po.lib().orc_kat_philox(_p(c), _p(k), _p(o))
    assert [int(v) for v in o] == exp


def test_normalize_and_systematic_and_weighted_index():
    L = po.lib()
    w = np.array([-1.0, -2.0, -3.0, -0.5])
    ref = np.exp(w - w.max())
    ref /= ref.sum()
    L.orc_kat_normalize(_p(w), 4)
    assert np.allclose(w, ref, rtol=1e-15)
    anc = np.zeros(4, dtype=np.uint64)
    wn = np.array([0.1, 0.2, 0.3, 0.4])
    L.orc_kat_systematic(0.5, _p(wn), 4, _p(anc))
    # targets .125 .375 .625 .875 against cumsum .1 .3 .6 1.0
    assert list(anc) == [1, 2, 3, 3]
    L.orc_kat_systematic(0.0, _p(wn), 4, _p(anc))  # first target 0: cum<target false -> ancestor 0
    assert list(anc) == [0, 1, 2, 3]
    assert L.orc_kat_weighted_index(0.0, _p(wn), 4) == 0
    assert L.orc_kat_weighted_index(0.1, _p(wn), 4) == 1  # cumulative weight <= chosen moves on
    assert L.orc_kat_weighted_index(0.999, _p(wn), 4) == 3


def test_sample_feature_from_probs_uses_raw_weights():  # Q3: cumsum passed, consumed as weights
    L = po.lib()
    probs = np.cumsum(np.ones(4))  # [1,2,3,4], total 10
    assert L.orc_kat_sample_feature(0.0, _p(probs), 4) == 0
    assert L.orc_kat_sample_feature(0.05, _p(probs), 4) == 0
    assert L.orc_kat_sample_feature(0.15, _p(probs), 4) == 1
    assert L.orc_kat_sample_feature(0.35, _p(probs), 4) == 2
    assert L.orc_kat_sample_feature(0.99, _p(probs), 4) == 3


def test_loglik_formulas():
    L = po.lib()
    rng = np.random.default_rng(1)
    y = rng.standard_normal(50)
    mu = rng.standard_normal(50)
    g = L.orc_kat_loglik(po.FAM_GAUSSIAN, 0.7, _p(y), _p(mu), 50)
    ref = np.sum(-0.5 * ((y - mu) / 0.7) ** 2 - np.log(0.7) - 0.5 * np.log(2 * np.pi))
    assert abs(g - ref) < 1e-10
    yb = (rng.random(50) < 0.5).astype(np.float64)
    b = L.orc_kat_loglik(po.FAM_BERNOULLI_LOGIT, 1.0, _p(yb), _p(mu), 50)
    refb = np.sum(yb * mu - np.logaddexp(0, mu))
    assert abs(b - refb) < 1e-10
    k = L.orc_kat_loglik(po.FAM_GAUSSIAN_KERNEL, 1.0, _p(y), _p(mu), 50)  # benches/particle_gibbs.rs:22-32
    assert abs(k - np.sum(-0.5 * (mu - y) ** 2)) < 1e-10
