"""Host logic of observation sharding on CPU (gloo, two ranks): row ranges and the global summaries every rank must agree on."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from bart_rs_b200.sharded import row_range


def test_row_ranges_cover_everything_once():
    for n in (1, 2, 7, 1000, 1000003):
        for world in (1, 2, 3, 8):
            rows = [row_range(n, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in rows]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bart_rs_b200.sharded import global_summaries
        rng = np.random.default_rng(3)
        X = rng.random((1001, 5), dtype=np.float32)
        y = rng.standard_normal(1001)
        lo, hi = row_range(1001, rank, world)
        q.put((rank,) + tuple(global_summaries(X[lo:hi], y[lo:hi])))
    finally:
        dist.destroy_process_group()


def test_global_summaries_agree_across_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    X = rng.random((1001, 5), dtype=np.float32)
    y = rng.standard_normal(1001)
    (_, n0, m0, lo0, hi0), (_, n1, m1, lo1, hi1) = out
    assert n0 == n1 == 1001
    assert m0 == m1                                      # bit-identical on both ranks
    assert abs(m0 - y.mean()) < 1e-12
    np.testing.assert_array_equal(lo0, lo1); np.testing.assert_array_equal(hi0, hi1)
    np.testing.assert_array_equal(lo0, X.min(axis=0)); np.testing.assert_array_equal(hi0, X.max(axis=0))
