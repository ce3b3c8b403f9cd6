"""Summarise an ncu --set full report into profiles/<name>_summary.csv (key metrics only) and, with --traffic KEY BATCH,
record its per-launch DRAM bytes in profiles/traffic.json (read by bench.py's roofline.traffic).
usage: python scripts/ncu_summary.py gpurun_out/r01b_fine_tokens_cl.ncu-rep [--traffic fine_tokens/cl 4]"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
name = os.path.splitext(os.path.basename(rep))[0] + (("_" + sys.argv[sys.argv.index("--suffix") + 1]) if "--suffix" in sys.argv else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
idx = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 0
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
KEYS = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
path = os.path.join(ROOT, "profiles", name + "_summary.csv")
with open(path, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i], vals[i]])
print("wrote", path)


def get(nm):
    i = hdr.index(nm)
    return float(vals[i]) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1)


if "--traffic" in sys.argv:
    j = sys.argv.index("--traffic")
    key, batch = sys.argv[j + 1], int(sys.argv[j + 2])
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    d = json.load(open(tpath)) if os.path.exists(tpath) else {}
    ti = hdr.index("gpu__time_duration.sum")
    dur = float(vals[ti]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(units[ti].lower(), 1)
    d[key] = {"batch": batch, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
              "kernel": vals[hdr.index("Kernel Name")], "duration_us_under_ncu": dur, "source": os.path.basename(rep)}
    json.dump(d, open(tpath, "w"), indent=1)
    print(key, d[key])
