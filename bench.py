#!/usr/bin/env python
"""bench.py — PG sweeps/sec of the B200-native PGBART sweep (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port)

A "step" is one full PG sweep: m = 200 consecutive tree updates = one PySampler.step() with
batch = (1.0, 1.0) (reference kernel.rs:60-126, pgbart.py:65).  Workload at N = 1 is BASELINE config C3
(synthetic Friedman n = 1M, p = 50, m = 200, 20 particles, Gaussian sigma = 1).  At N > 1 (torchrun, one rank per
GPU) every rank runs an independent chain of the same workload with its own seed — the "chains" sharding of
SURVEY.md §8(e): no data-path collective, weak scaling; value = sweeps of all ranks / max-over-ranks time.

Every run also carries, in the same JSON line, a `sharded` block: workload C4 (n = 10M, p = 100) as ONE chain whose
rows are split over the N GPUs (observation sharding; at N = 1 the same workload on one GPU), with
ms per sweep, sweeps/s and rank 0's roofline figures, plus `sharded_parity_ok` from running the smallest case of
tests/run_sharded.py against the oracle before timing.  `value` stays the chains number so it is comparable across N.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant (and only) hot kernel pg_sweep_persistent: achieved = SURVEY.md §8(d) algorithmic bytes of the
                sweep (device counters) / CUDA-event time on the sampler's stream; fused_* = bytes this
                implementation must move; traffic = ncu dram bytes (profiles/ncu_traffic.json); peak = MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++ restatement of the bart-rs CPU path, 1 core) timed here on a bounded sample
  e2e           same metric through PySampler.step() with host buffers (likelihood push H2D, predictions D2H)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": dict(n=1_000, p=10, m=50, P=20, family="gaussian"),
    "C2": dict(n=100_000, p=20, m=200, P=20, family="gaussian"),
    "C3": dict(n=1_000_000, p=50, m=200, P=20, family="gaussian"),
    "C5": dict(n=1_000_000, p=50, m=200, P=256, family="bernoulli_logit"),
    # observation-sharded: ONE chain, rows split across the GPUs; strong scaling
    "C4": dict(n=10_000_000, p=100, m=200, P=20, family="gaussian", sharded=True),
    "C3s": dict(n=1_000_000, p=50, m=200, P=20, family="gaussian", sharded=True),
    "C3x20": dict(n=20_000_000, p=50, m=200, P=20, family="gaussian"),     # how close to the HBM roofline large n gets
}


def make_data(w, seed=0, rows=None):
    """rows = (lo, hi): only this rank's rows of a sharded workload (generated from a per-shard seed; the leaf sd is
    derived from the local rows, which is the same number to three digits and only needs to agree across ranks)."""
    from tests import datagen
    n = w["n"] if rows is None else rows[1] - rows[0]
    X, y = datagen.friedman(n, w["p"], seed=seed if rows is None else seed + 1000 + rows[0] % 9973,
                            bernoulli=(w["family"] == "bernoulli_logit"))
    prm = datagen.pgbart_params(y, w["m"], w["p"])
    return X, y, prm


def config_for(name, world):
    """The `config` object of the JSON line: both arms (this repo's and --impl reference) print the same one."""
    w = WORKLOADS[name]
    sharded = bool(w.get("sharded")) and world > 1
    rows = w["n"] // world if sharded else w["n"]
    x_mb = rows * w["p"] * 4 / 1e6
    return {"workload": f"{name}: Friedman n={w['n']} p={w['p']} m={w['m']} P={w['P']} {w['family']}, "
                        "batch=(1,1) so one step = one sweep",
            "parallelism": "1 GPU" if world == 1 else (f"one chain, rows sharded over {world} GPUs, in-kernel exchange of sufficient "
                                                         "statistics over NVLink peer memory" if sharded
                                                         else f"{world} independent chains, one per GPU, no collective"),
            "l2": f"inputs larger than L2 (X {x_mb:.0f} MB fp32 per GPU vs 126 MB L2); no flush" if x_mb * 1e6 > 126e6
                  else "working set fits L2 (L2-resident by design of this config)",
            "storage": "X, y fp32; running sum-of-trees fp64; accumulation fp64"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bounded_updates(w, updates, budget_s=15.0):
    """Tree updates of the CPU sample, cut so that the sample stays near `budget_s` seconds: the oracle (like the
    reference) costs about 1.8e-8 s (Gaussian) / 3.5e-8 s (Bernoulli) per row per growing particle per tree update."""
    per = w["n"] * max(w["P"] - 1, 1) * (3.5e-8 if w["family"] == "bernoulli_logit" else 1.8e-8)
    return int(min(max(updates, 1), max(2, int(budget_s / per))))


def oracle_sample(X, y, prm, w, updates, seed=0, warm=2):
    """Time `updates` tree updates of the CPU oracle on the workload (single core, like the reference)."""
    from oracle import pyoracle as po
    m = w["m"]
    updates = bounded_updates(w, updates)
    warm = min(warm, updates)
    frac = max(updates, 1) / m
    o = po.OracleSampler(X.astype(np.float64), y, n_trees=m, n_particles=w["P"], max_depth=prm["max_depth"],
                         alpha=0.95, beta=2.0, sigma=prm["sigma"], split_prior=prm["split_prior"],
                         batch_tune=min(warm / m, 1.0), batch_post=min(frac, 1.0), seed=seed)
    o.set_likelihood(po.FAM_BERNOULLI_LOGIT if w["family"] == "bernoulli_logit" else po.FAM_GAUSSIAN, [1.0])
    o.set_rng(po.RNG_PHILOX, seed)
    o.step(True)                      # warm-up: page in, allocate
    t0 = time.perf_counter()
    o.step(False)
    dt = time.perf_counter() - t0
    done = min(max(int(np.floor(frac * m + 0.5)), 1), m)
    o.close()
    return done, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be built in
    this image (no cargo/rustc), so this is the C++ oracle port, single-threaded like the reference
    (`#[pyclass(unsendable)]`, no threads in src/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    X, y, prm = make_data(w)
    per_step = bounded_updates(w, max(1, int(args.ref_updates)))
    # the whole --steps K --warmup W run should end within a few minutes: about two minutes of CPU work in total
    per_update_s = w["n"] * max(w["P"] - 1, 1) * (3.5e-8 if w["family"] == "bernoulli_logit" else 1.8e-8)
    per_step = max(min(per_step, int(120.0 / ((args.warmup + args.steps) * per_update_s))), min(per_step, 4))
    times, done_total = [], 0
    for i in range(args.warmup + args.steps):
        done, dt = oracle_sample(X, y, prm, w, per_step, seed=i, warm=1)
        if i >= args.warmup:
            times.append(dt)
            done_total += done
    total = float(np.sum(times))
    sweeps_per_s = (done_total / w["m"]) / total
    line = {
        "impl": "reference", "metric": "PG sweeps/sec", "value": sweeps_per_s, "unit": "sweeps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args.workload, args.gpus),
        "cpu_baseline": {"value": sweeps_per_s, "unit": "sweeps/s", "cores": 1, "kind": "port",
                         "sample": f"each step = {per_step} tree updates of the sweep on 1 host core (the reference is single-threaded); "
                                   f"{done_total} tree updates in {total:.2f} s, scaled to m={w['m']}"},
        "e2e": {"value": sweeps_per_s, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_sweeps(smp, steps, warmup, barrier):
    """K sweeps back to back on the sampler's stream, CUDA events around them; returns (ms_total, counters before / after)."""
    for _ in range(warmup):
        smp.step_device(False)
    smp.sync()
    c0 = smp.counters()
    barrier()
    smp.timer_start()
    for _ in range(steps):
        smp.step_device(False)
    ms = smp.timer_stop()
    smp.sync()
    barrier()
    return ms, c0, smp.counters()


def sharded_block(args, rank, local_rank, world, barrier, peak):
    """Workload C4 as one chain over `world` GPUs (rows sharded; world = 1: the whole workload on one GPU)."""
    import torch
    import torch.distributed as dist
    from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler
    w = WORKLOADS[args.sharded_workload]
    parity_ok = None
    if world > 1:
        from tests import run_sharded
        run_sharded.QUIET = True
        ok = run_sharded.run_case(run_sharded.CASES[0], rank, local_rank, world)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        parity_ok = bool(int(flag.item()))
    lik = DeviceLikelihood(w["family"], 1.0)
    if world > 1:
        from bart_rs_b200.sharded import ShardedSampler, row_range
        lo, hi = row_range(w["n"], rank, world)
        X, y, prm = make_data(w, rows=(lo, hi))
        sig = torch.tensor([prm["sigma"]], dtype=torch.float64, device="cuda")
        dist.broadcast(sig, 0)
        st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, float(sig.item()), list(prm["split_prior"]),
                            ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
        smp = ShardedSampler.init(X, y, lik, st, device=local_rank)
    else:
        X, y, prm = make_data(w)
        st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                            ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
        smp = PySampler.init(X, y, lik, st, device=local_rank)
    rows_local = X.shape[0]
    del X
    steps = max(2, min(args.steps, args.sharded_steps))
    ms, c0, c1 = time_sweeps(smp, steps, 3, barrier)
    from bart_rs_b200.chains import job_throughput
    _, ms_max = job_throughput(ms, steps, world, device="cuda")
    # end to end: this rank's rows of the predictions come back to the host every step
    try:
        smp.step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            smp.set_likelihood(lik)
            smp.step(False)
        torch.cuda.synchronize()
    except RuntimeError as e:
        raise RuntimeError(f"{e}; rank {rank} device state {smp.debug_state()}") from None
    e2e_ms = 1000.0 * (time.perf_counter() - t0) / 2
    _, e2e_ms_max = job_throughput(e2e_ms, 1, world, device="cuda")
    smp.close()
    bm = (c1["bytes_model"] - c0["bytes_model"]) / steps
    bc = (c1["bytes_contract"] - c0["bytes_contract"]) / steps
    sec = ms / 1000.0 / steps
    return {
        "workload": f"{args.sharded_workload}: Friedman n={w['n']} p={w['p']} m={w['m']} P={w['P']} {w['family']}, one chain",
        "n_gpus": world, "rows_per_gpu": rows_local, "steps": steps, "scaling": "strong",
        "parallelism": "1 GPU" if world == 1 else f"rows sharded over {world} GPUs; per-proposal sufficient statistics exchanged inside "
                                                  "the persistent kernel through NVLink peer mailboxes, combined in rank order",
        "ms_per_sweep": ms_max / steps, "sweeps_per_s": steps / (ms_max / 1000.0),
        "e2e_ms_per_sweep": e2e_ms_max, "e2e_sweeps_per_s": 1000.0 / e2e_ms_max,
        "roofline_rank0": {"algorithmic_bytes_per_launch": bc, "achieved": bc / sec / 1e9, "frac": bc / sec / 1e9 / peak,
                           "fused_bytes_per_launch": bm, "fused_gbs": bm / sec / 1e9, "fused_frac": bm / sec / 1e9 / peak,
                           "unit": "GB/s", "peak": peak, "note": "bytes of rank 0's own rows / its own launch time"},
        "parity_ok": parity_ok,
        "parity_case": None if world == 1 else "tests/run_sharded.py CASES[0] (n=5003, p=8, m=12, P=20, 2 sweeps) vs the oracle on all rows",
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this repo has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from bart_rs_b200 import DeviceLikelihood, PyBartSettings, PySampler

    w = WORKLOADS[args.workload]
    from bart_rs_b200.chains import chain_settings, job_throughput
    sharded = bool(w.get("sharded")) and world > 1
    lik = DeviceLikelihood(w["family"], 1.0)
    if sharded:
        # one chain over all rows: rank r generates and keeps rows [lo, hi); sigma must be the same number on every rank
        from bart_rs_b200.sharded import ShardedSampler, row_range
        lo, hi = row_range(w["n"], rank, world)
        X, y, prm = make_data(w, rows=(lo, hi))
        sig = torch.tensor([prm["sigma"]], dtype=torch.float64, device="cuda")
        dist.broadcast(sig, 0)
        st = PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, float(sig.item()), list(prm["split_prior"]),
                            ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0)
        smp = ShardedSampler.init(X, y, lik, st, device=local_rank)
    else:
        X, y, prm = make_data(w)
        st = chain_settings(PyBartSettings(w["m"], w["P"], prm["max_depth"], 0.95, 2.0, prm["sigma"], list(prm["split_prior"]),
                                           ["ContinuousSplit"] * w["p"], "constant", "systematic", 1.0, 1.0, seed=0), rank)
        smp = PySampler.init(X, y, lik, st, device=local_rank)   # X is float32 C-order: converted to column-major on upload

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs already in HBM, K sweeps, CUDA events on the sampler's stream
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, c0, c1 = time_sweeps(smp, args.steps, args.warmup, barrier)
    clk = clocks.stop() if rank == 0 else None
    value, ms_max = job_throughput(ms, args.steps, world, device="cuda")
    if sharded:
        value /= world            # all ranks worked on the SAME sweeps

    # ---- end to end: PySampler.step() with host buffers (likelihood parameters H2D, predictions D2H each step)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    smp.set_likelihood(lik)
    smp.step(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        smp.set_likelihood(lik)          # the reference refreshes shared likelihood inputs before every step (pgbart.py:190)
        pred = smp.step(False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value, _ = job_throughput(1000.0 * e2e_s, e2e_steps, world, device="cuda")
    if sharded:
        e2e_value /= world
    assert np.all(np.isfinite(pred))

    # ---- reported beside the headline, never instead of it: the same sweep with option prune_dead_proposals
    # (proposals of particles whose weight is provably exactly 0 are not evaluated; identical forest and predictions)
    pruned = None
    if world == 1 and w["family"] == "gaussian":
        smp.set_option("prune_dead_proposals", "1")
        for _ in range(3):
            smp.step_device(False)
        smp.sync()
        smp.timer_start()
        for _ in range(args.steps):
            smp.step_device(False)
        pms = smp.timer_stop()
        smp.sync()
        smp.set_option("prune_dead_proposals", "0")
        pruned = {"value": args.steps / (pms / 1000.0), "unit": "sweeps/s", "ms_per_step": pms / args.steps,
                  "note": "option prune_dead_proposals=1 (off by default; DESIGN.md section 4b): same forest and predictions, "
                          "dead proposals not evaluated; NOT the headline value"}

    spec = smp.spec_counters()
    smp.close()
    shard = None
    if not args.no_sharded and not w.get("sharded"):
        try:
            shard = sharded_block(args, rank, local_rank, world, barrier, peaks()[0])
        except RuntimeError as e:
            # a device-side failure of the sharded run (e.g. an exchange that timed out) is reported by every rank in an
            # orderly way (pgbart_step returns the error): the headline line is still printed, with the error in the block
            shard = {"workload": WORKLOADS[args.sharded_workload]["n"], "n_gpus": world, "error": str(e)[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        k = args.steps
        # which persistent kernel ran (pg_launch_persistent, csrc/pg_sweep.cu): one per pass variant
        if w["P"] > 32:
            kernel_name = "pg_sweep_persistent_generic"
        elif w["family"] == "bernoulli_logit":
            kernel_name = "pg_sweep_persistent_bern"
        else:
            kernel_name = "pg_sweep_persistent_ring" if X.shape[0] <= c1["grid_blocks"] * 2048 else "pg_sweep_persistent_direct"
        bytes_per_launch = (c1["bytes_model"] - c0["bytes_model"]) / k
        contract_per_launch = (c1["bytes_contract"] - c0["bytes_contract"]) / k
        fused_gbs = bytes_per_launch / (ms / 1000.0 / k) / 1e9
        achieved = contract_per_launch / (ms / 1000.0 / k) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("workload") == args.workload:
                traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            done, dt = oracle_sample(X, y, prm, w, args.ref_updates)
            cpu = {"value": (done / w["m"]) / dt, "unit": "sweeps/s", "cores": 1, "kind": "port",
                   "sample": f"{done} tree updates of this workload in {dt:.2f} s on 1 host core, scaled to m={w['m']} "
                             "(C++ restatement of the bart-rs CPU path; the Rust crate cannot be built in this image)"}
        line = {
            "metric": "PG sweeps/sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.workload, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": kernel_name + " (one launch = one sweep = m tree updates)"
                                   + (f"; rank 0 of {world}: bytes of its own rows, per-GPU figure" if sharded else ""),
                         "algorithmic_bytes_per_launch": contract_per_launch,
                         "fused_bytes_per_launch": bytes_per_launch, "fused_gbs": fused_gbs, "fused_frac": fused_gbs / peak,
                         "note": "achieved = SURVEY.md 8(d) contract bytes (B_tree summed from device counters over the sweep's "
                                 "grows: per-slot min/max + stats + partition passes, fp64 A, no credit for cross-particle "
                                 "sharing) / CUDA-event time of the launch; it can exceed the HBM peak because the formula charges A, y and the "
                                 "index list once per slot while the fused root pass reads them once per tree.  fused_* = the bytes this implementation must move "
                                 "(A, y and each distinct split column read once per pass for all slots together); "
                                 "traffic = ncu dram bytes of one launch (A and y stay L2-resident, so it is below both). "
                                 "The launch is bound by latency (control sections and per-pass instruction issue), not by HBM: DESIGN.md section 4-5"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8 * X.shape[0],
                    "steps": e2e_steps},
            "pruned": pruned,
            "sharded": shard,
            "sharded_parity_ok": None if shard is None else shard.get("parity_ok"),
            "speculation": spec,
            "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
            "grid_passes_per_step": (c1["grid_passes"] - c0["grid_passes"]) / k,
            "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-updates", type=int, default=40, help="tree updates per timed CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded-workload", default="C4", choices=["C4", "C3s"], help="workload of the `sharded` block")
    ap.add_argument("--sharded-steps", type=int, default=5)
    ap.add_argument("--no-sharded", action="store_true", help="skip the observation-sharded block")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
