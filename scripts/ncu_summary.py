"""Compact summary of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py > out.csv`
keeps, per profiled kernel, the metrics the roofline discussion in DESIGN.md refers to."""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
w = csv.writer(sys.stdout)
w.writerow(KEEP + ["top_stalls (warps stalled per issue-active cycle)"])
w.writerow([units[idx[k]] if k in idx else "" for k in KEEP] + [""])
for d in data:
    top = sorted(((float(d[idx[h]]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h in stalls), reverse=True)[:4]
    w.writerow([d[idx[k]] if k in idx else "" for k in KEEP] + ["; ".join(f"{n}={v:.2f}" for v, n in top)])
