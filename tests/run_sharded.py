"""Observation-sharded parity run: launch with torchrun (one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_sharded.py

Every rank builds the same synthetic data, keeps its own rows, and runs a ShardedSampler with the free-running Philox
generator.  Rank 0 runs the CPU oracle on ALL rows with the same generator and checks: forest structure and split values
bit-exact (every rank holds the whole forest), leaf values / predictions / log-likelihood within 1e-6, BartInfo equal.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bart_rs_b200 import DeviceLikelihood, PyBartSettings          # noqa: E402
from bart_rs_b200.sharded import ShardedSampler, row_range           # noqa: E402
from tests import datagen, parity                                    # noqa: E402

CASES = [
    # n, p, m, P, family, steps
    (5003, 8, 12, 20, "gaussian", [True, False]),
    (40000, 10, 10, 20, "gaussian", [True, True]),
    (3001, 4, 6, 3, "gaussian", [True, True, True]),       # few particles: deep trees, partitions of non-root nodes
    (4000, 5, 6, 6, "bernoulli_logit", [True, False]),
    (200000, 20, 16, 20, "gaussian", [True, True]),        # larger shards: fp64 summation order differs most from one GPU
]


def run_case(case, rank, local, world):
    """One observation-sharded run checked against the oracle on ALL rows (rank 0 checks; returns its verdict on every rank
    only after the caller broadcasts it).  The process group must exist.  Also used by bench.py for `sharded_parity_ok`."""
    (n, p, m, P, family, tunes) = case
    bern = family == "bernoulli_logit"
    X, y = datagen.small(n, p, seed=n + m, bernoulli=bern)
    kw = dict(beta=0.7, max_depth=10) if P == 3 else {}
    prm = datagen.pgbart_params(y, m, p, 0.95, kw.get("beta", 2.0))
    max_depth = kw.get("max_depth", prm["max_depth"])
    st = PyBartSettings(m, P, max_depth, 0.95, kw.get("beta", 2.0), prm["sigma"], list(prm["split_prior"]),
                        ["ContinuousSplit"] * p, "constant", "systematic", 1.0, 1.0, seed=7)
    lo, hi = row_range(n, rank, world)
    smp = ShardedSampler.init(X[lo:hi], y[lo:hi], DeviceLikelihood(family, 1.0), st, device=local)
    smp.set_rng_philox(7)
    preds = [smp.step(t) for t in tunes]
    gathered = [None] * world
    dist.all_gather_object(gathered, [np.asarray(q) for q in preds])
    ok = True
    if rank == 0:
        from oracle import pyoracle as po
        orc = po.OracleSampler(X.astype(np.float64), y, n_trees=m, n_particles=P, max_depth=max_depth, alpha=0.95,
                               beta=kw.get("beta", 2.0), sigma=prm["sigma"], split_prior=prm["split_prior"],
                               batch_tune=1.0, batch_post=1.0, seed=7)
        orc.set_likelihood(po.FAM_BERNOULLI_LOGIT if bern else po.FAM_GAUSSIAN, [1.0])
        orc.set_rng(po.RNG_PHILOX, 7)
        try:
            for k, t in enumerate(tunes):
                ref = np.asarray(orc.step(t))
                got = np.concatenate([g[k] for g in gathered])
                np.testing.assert_allclose(got, ref, rtol=parity.RTOL, atol=0, err_msg=f"predictions step {k}")
            parity.assert_forest_matches(orc, smp, m, check_leaf_indices=False)
            oll, oacc, odep = orc.info()
            dll, dacc, ddep = smp.info()
            assert oacc == dacc, (oacc, dacc)
            np.testing.assert_array_equal(ddep, odep)
            np.testing.assert_allclose(dll, oll, rtol=1e-5, atol=1e-12)
            print(f"sharded parity OK: n={n} p={p} m={m} P={P} {family} world={world}", file=sys.stderr if QUIET else sys.stdout, flush=True)
        except AssertionError as e:
            ok = False
            print(f"sharded parity FAILED: n={n} p={p} m={m} P={P} {family}: {e}", file=sys.stderr if QUIET else sys.stdout, flush=True)
    smp.close()
    dist.barrier()
    return ok


QUIET = False      # bench.py sets this: its stdout carries exactly one JSON line


def main():
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for case in CASES:
        ok = run_case(case, rank, local, world) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
