"""Phase breakdown of an observation-sharded sweep on rank 0 (torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/profile_sharded.py [workload] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from bart_rs_b200 import DeviceLikelihood, PyBartSettings
from bart_rs_b200.sharded import ShardedSampler, row_range

wl = sys.argv[1] if len(sys.argv) > 1 else "C3s"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
w = bench.WORKLOADS[wl]
lo, hi = row_range(w["n"], rank, world)
X, y, prm = bench.make_data(w, rows=(lo, hi))
sig = torch.tensor([prm["sigma"]], dtype=torch.float64, device="cuda")
dist.broadcast(sig, 0)
st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, float(sig.item()), list(prm["split_prior"]),
                    ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
s = ShardedSampler.init(X, y, DeviceLikelihood(w["family"], 1.0), st, device=local)
torch.cuda.synchronize()
dist.barrier()
try:
    for i in range(2):
        s.step_device(False)
        s.sync()
except RuntimeError as e:
    print(f"[rank {rank}] FAILED in warm-up step {i}: {e}; state {s.debug_state()}", flush=True)
    os._exit(3)
dist.barrier()
p0 = s.profile()
s.timer_start()
for _ in range(steps):
    s.step_device(False)
ms = s.timer_stop()
s.sync()
p1 = s.profile()
t = torch.tensor([ms / steps], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ghz = 1.965
    print(f"{wl} on {world} GPUs: {float(t.item()):.3f} ms per sweep (max over ranks); rank 0 rows {hi - lo}")
    for n in ("root", "stats", "part"):
        cnt = (p1["passes"][n] - p0["passes"][n]) / steps
        a, b, c = ((p1[k][n] - p0[k][n]) / steps / max(cnt, 1) / ghz / 1e3 for k in ("pass_cycles", "wait_serial_cycles", "serial_cycles"))
        print(f"{n:6s} passes/sweep {cnt:6.1f}  pass {a:7.2f} us  wait+serial {b:7.2f} us  serial {c:7.2f} us")
    for n in p1["serial_sub_cycles"]:
        cyc = (p1["serial_sub_cycles"][n] - p0["serial_sub_cycles"][n]) / steps
        calls = (p1["serial_sub_calls"][n] - p0["serial_sub_calls"][n]) / steps
        if calls:
            print(f"  serial {n:14s} calls/sweep {calls:7.1f}  us/call {cyc / calls / ghz / 1e3:7.2f}  ms/sweep {cyc / ghz / 1e6:6.2f}")
    ctl = {k: (p1["control"][k] - p0["control"][k]) / steps for k in p1["control"]}
    print("control idle (ms/sweep): " + ", ".join(f"{k[5:]} {ctl[k] / ghz / 1e6:.2f}" for k in ("wait_root", "wait_stats", "wait_part", "wait_other")),
          "; speculative root passes per sweep", (p1["speculative_root_passes"] - p0["speculative_root_passes"]) / steps)
s.close()
dist.destroy_process_group()
