"""Independent chains across GPUs — the only multi-GPU sharding of this path that needs no exchange step.

SURVEY.md §8(e): trees are sequentially dependent through the running predictions and every particle needs
the same rows, so one chain does not split by tree or particle; PyMC runs one sampler per chain
(reference python/pymc_bart/pgbart.py:61-185 is instantiated once per chain by pm.sample, seed = chain id).
Here: one process per GPU (torchrun), rank r runs chain r with seed base+r, and the only communication is the
max-over-ranks of the timed region for reporting.  No data-path collective exists or is invented.
"""
from __future__ import annotations

import copy


def chain_seed(base_seed: int, rank: int) -> int:
    """Seed of the chain owned by `rank` (chain id added to the user's seed, like PyMC's per-chain seeds)."""
    return int(base_seed) + int(rank)


def chain_settings(settings, rank: int):
    """Copy of a PyBartSettings for the chain owned by `rank`."""
    out = copy.copy(settings)
    out.seed = chain_seed(settings.seed, rank)
    return out


def job_throughput(local_ms: float, steps_per_rank: int, world: int, device=None):
    """Whole-job sweeps/s: sweeps of all ranks / max-over-ranks time.  Returns (value, ms_max).

    Uses torch.distributed when world > 1 (NCCL on GPUs, gloo in the CPU tests); the reduction carries one
    float64 and is outside every timed region."""
    ms_max = float(local_ms)
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([local_ms], dtype=torch.float64, device=device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    return world * steps_per_rank / (ms_max / 1000.0), ms_max
