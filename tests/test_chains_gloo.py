"""N > 1 path on CPU: two gloo ranks, each owning one independent chain (SURVEY.md §8(e) "chains").

The host logic under test is what bench.py and a multi-GPU user run around the sampler: per-rank chain seeds, no
data-path collective, and the max-over-ranks time reduction.  The chain itself is stood in for by the CPU oracle
(test infrastructure) because there is no GPU here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import datagen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bart_rs_b200 import PyBartSettings
        from bart_rs_b200.chains import chain_settings, job_throughput
        from oracle import pyoracle as po
        X, y = datagen.small(300, 4, seed=5)
        prm = datagen.pgbart_params(y, 8, 4)
        base = PyBartSettings(8, 6, prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                              ["ContinuousSplit"] * 4, "constant", "systematic", 1.0, 1.0, seed=3)
        st = chain_settings(base, rank)
        o = po.OracleSampler(X.astype(np.float64), y, n_trees=st.n_trees, n_particles=st.n_particles,
                             max_depth=st.max_depth, alpha=st.alpha, beta=st.beta, sigma=st.sigma,
                             split_prior=st.split_prior, batch_tune=1.0, batch_post=1.0, seed=st.seed)
        o.set_likelihood(po.FAM_GAUSSIAN, [1.0])
        o.set_rng(po.RNG_PHILOX, st.seed)
        pred = np.array(o.step(False), dtype=np.float64)
        o.close()
        local_ms = 10.0 * (rank + 1)                       # rank 1 is the slow one
        value, ms_max = job_throughput(local_ms, 2, world)
        gathered = [torch.zeros(300, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(pred))  # test-only exchange, to compare the chains
        q.put((rank, st.seed, value, ms_max, [g.numpy().copy() for g in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_chains_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, s0, v0, m0, g0), (r1, s1, v1, m1, g1) = out
    assert (s0, s1) == (3, 4)                              # chain id added to the user's seed
    assert m0 == m1 == 20.0                                # max over ranks, seen identically by both
    assert v0 == v1 == pytest.approx(2 * 2 / 0.020)        # sweeps of all ranks / slowest rank's time
    assert np.array_equal(g0[0], g1[0]) and np.array_equal(g0[1], g1[1])
    assert not np.array_equal(g0[0], g0[1])                # different seeds -> different chains
    assert np.all(np.isfinite(g0[0])) and np.all(np.isfinite(g0[1]))


def test_single_rank_needs_no_process_group():
    from bart_rs_b200.chains import job_throughput
    v, ms = job_throughput(50.0, 5, 1)
    assert ms == 50.0 and v == pytest.approx(100.0)
