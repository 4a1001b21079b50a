"""Key metrics of one kernel from an .ncu-rep: python scratch/ncu_summary.py rep kernel-substring"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, u = r[0], r[1]
ki = h.index("Kernel Name")
rows = [x for x in r[2:] if pat in x[ki]]
row = rows[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
print(row[ki][:110])
for w in want:
    if w in h:
        i = h.index(w)
        print("  %-70s %s %s" % (w, row[i], u[i]))
st = [(h[i], row[i]) for i in range(len(h)) if h[i].startswith("smsp__average_warps_issue_stalled") and h[i].endswith("per_issue_active.ratio")]
for n, v in sorted(st, key=lambda t: -float(t[1] or 0))[:8]:
    print("  stall %-55s %s" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
