"""Observation sharding (BASELINE.json north_star): one chain, rows split across the GPUs of a node.

One process per GPU (torchrun); rank r owns the contiguous rows [r n / W, (r + 1) n / W).  torch.distributed is the
plumbing only: it carries the global summary statistics the sampler is created with and the CUDA IPC handles of the
per-rank mailboxes.  The exchange of per-proposal sufficient statistics itself happens inside the persistent kernel
(bart_rs_b200/csrc/pg_serial_fast.cuh, fw_exchange): stores into peer memory over NVLink, combined in rank order, so
every rank takes bit-identical decisions — no collective is launched per reduction.

The reference has no counterpart (it is a single-process CPU sampler); the per-rank object mirrors PySampler.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .pymc_bart import PySampler


def row_range(n: int, rank: int, world: int):
    """Rows owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_summaries(x_local: np.ndarray, y_local: np.ndarray, group=None):
    """(n_global, mean of all y, per-column global min, max) via torch.distributed (any backend with CPU tensors or, for
    NCCL, CUDA tensors).  The mean is sum / n over ranks in rank order of the partial sums — all ranks get the same bits."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    stats = torch.tensor([float(x_local.shape[0]), float(np.sum(y_local, dtype=np.float64))], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats, group=group)
    n_global = int(sum(float(g[0]) for g in gathered))
    y_sum = 0.0
    for g in gathered:                      # fixed order: identical on every rank
        y_sum += float(g[1])
    cmin = torch.from_numpy(np.min(x_local, axis=0).astype(np.float32)).to(dev)
    cmax = torch.from_numpy(np.max(x_local, axis=0).astype(np.float32)).to(dev)
    dist.all_reduce(cmin, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(cmax, op=dist.ReduceOp.MAX, group=group)
    return n_global, y_sum / n_global, cmin.cpu().numpy(), cmax.cpu().numpy()


class ShardedSampler(PySampler):
    """PySampler over this rank's rows of an observation-sharded chain.  `step()` returns this rank's predictions."""

    @staticmethod
    def init(x_local, y_local, model, settings, *, device: int, group=None) -> "ShardedSampler":
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        n_global, y_mean, cmin, cmax = global_summaries(np.asarray(x_local), np.asarray(y_local), group)
        base = PySampler.init(x_local, y_local, model, settings, device=device)
        self = ShardedSampler.__new__(ShardedSampler)
        self.__dict__.update(base.__dict__)
        base._h = None                      # ownership moved
        L = self._L
        cmin = np.ascontiguousarray(cmin, dtype=np.float32)
        cmax = np.ascontiguousarray(cmax, dtype=np.float32)
        self._check(L.pgbart_shard_config(self._h, rank, world, n_global, float(y_mean),
                                          cmin.ctypes.data_as(C.c_void_p), cmax.ctypes.data_as(C.c_void_p)))
        handle = np.zeros(64, dtype=np.uint8)
        self._check(L.pgbart_shard_mailbox(self._h, handle.ctypes.data_as(C.c_void_p)))
        handles = [None] * world
        dist.all_gather_object(handles, handle.tobytes(), group=group)
        blob = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        self._check(L.pgbart_shard_connect(self._h, blob.ctypes.data_as(C.c_void_p)))
        # Nothing may be allocated or page-locked inside a step: a rank that stalls on the host while its peers' kernels
        # already wait for it in the exchange runs into their watchdog.  Pin the output blocks and the staging buffer now.
        self._pool.prewarm(4)
        self.set_option("prewarm_host", "1")
        # Every mailbox must be mapped everywhere before the first step, and the ranks' HOSTS must leave together: the
        # control blocks wait for each other inside the kernel (with a watchdog), while mapping seven peers' memory can
        # take a rank many seconds longer than another.  dist.barrier() on NCCL only orders streams, so gather an object
        # instead (that blocks the host on every backend).
        done = [None] * world
        dist.all_gather_object(done, rank, group=group)
        self.rank, self.world, self.n_global = rank, world, n_global
        return self
