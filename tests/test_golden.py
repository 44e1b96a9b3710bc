"""Golden fixtures (tests/golden): the oracle must reproduce the stored bytes; the host simulation of the device
control code and (on a GPU) the CUDA path must replay them to the stated tolerance."""
import numpy as np
import pytest

from tests import parity
from tests.golden import make_golden as mg

BIT_KEYS = ["tape_slot", "tape_round", "tape_upd", "trace_slot", "trace_round", "trace_upd", "n_upd", "preds",
            "split_var", "split_val", "leaf_val", "leaf_indices", "info_acc", "info_depths", "info_ll"]


def _rerun(g, replay):
    meta = g["meta"]
    cfg = {k: meta[k] for k in ("n_trees", "n_particles", "max_depth", "alpha", "beta", "sigma", "split_prior",
                                "batch_tune", "batch_post", "seed")}
    tape = (g["tape_slot"], g["tape_round"], g["tape_upd"]) if replay else None
    return mg.run_oracle(g["X"], g["y"], cfg, meta["family"], meta["lik_sigma"], meta["tunes"], tape=tape)


@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
def test_oracle_reproduces_golden_bytes(name):
    g = parity.load_golden(name)
    out = _rerun(g, replay=False)
    for k in BIT_KEYS:
        a, b = np.asarray(out[k]), np.asarray(g[k])
        assert a.shape == b.shape, k
        assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k


@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
def test_oracle_replay_of_golden_tape(name):
    g = parity.load_golden(name)
    out = _rerun(g, replay=True)
    for k in ["trace_slot", "trace_round", "trace_upd", "preds", "split_var", "leaf_val", "leaf_indices"]:
        assert np.array_equal(np.asarray(out[k]), np.asarray(g[k]), equal_nan=(np.asarray(g[k]).dtype.kind == "f")), k


@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
def test_host_simulation_replays_golden(name):
    from tests.test_sim_parity import make_sim
    parity.golden_parity(make_sim(3), parity.load_golden(name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", parity.GOLDEN_NAMES)
@pytest.mark.parametrize("driver", ["persistent", "stepwise"])
def test_gpu_replays_golden(name, driver):
    from tests.test_gpu_parity import make_gpu
    parity.golden_parity(make_gpu(driver), parity.load_golden(name))
