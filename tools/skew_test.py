"""Does a host-side allocation on one rank stall behind another rank's persistent kernel?  (torchrun, 2 GPUs)
Rank 1 sleeps before its step; mode "alloc" lets step() page-lock its output block at that moment, mode "warm" has pinned
the blocks beforehand.  python -m torch.distributed.run --nproc-per-node 2 ... tools/skew_test.py alloc|warm"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings
from bart_rs_b200.sharded import ShardedSampler, row_range

mode = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
w = dict(bench.WORKLOADS["C3s"]); w["n"] = 400_000
lo, hi = row_range(w["n"], rank, world)
X, y, prm = bench.make_data(w, rows=(lo, hi))
sig = torch.tensor([prm["sigma"]], dtype=torch.float64, device="cuda"); dist.broadcast(sig, 0)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, float(sig.item()), list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 0.1, 0.1, seed=0)
s = ShardedSampler.init(X, y, DeviceLikelihood("gaussian", 1.0), st, device=local)
s.step_device(False); s.sync()
if mode == "warm":
    a = s.step(False); b = s.step(False); del a, b
if rank == 1:
    time.sleep(3.0)
t0 = time.perf_counter()
try:
    out = s.step(False)
    print(f"[rank {rank}] mode {mode}: step with a {'fresh' if mode == 'alloc' else 'pre-pinned'} output block took {time.perf_counter() - t0:.2f} s", flush=True)
except RuntimeError as e:
    print(f"[rank {rank}] mode {mode}: FAILED after {time.perf_counter() - t0:.2f} s: {e}", flush=True)
    os._exit(0)
s.close()
dist.destroy_process_group()
