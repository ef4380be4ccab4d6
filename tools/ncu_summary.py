#!/usr/bin/env python
"""tools/ncu_summary.py report.ncu-rep — one row per captured launch with the counters the profiles/*.md tables use."""
import csv, subprocess, sys
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
     "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "smsp__thread_inst_executed_per_inst_executed.ratio",
     "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
     "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
     "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
     "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
     "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    print("==", d["ID"], name)
    for m in M:
        if m in d and d[m] != "":
            print(f"   {m:75s} {d[m]}")
    try:
        print(f"   lanes per instruction {float(d['smsp__thread_inst_executed.sum'].replace(',','')) / float(d['smsp__inst_executed.sum'].replace(',','')):.2f}")
    except Exception:
        pass
