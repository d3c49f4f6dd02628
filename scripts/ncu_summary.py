"""Print the handful of ncu raw metrics that matter here (development aid): python scripts/ncu_summary.py rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_barriers", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
            try:
                if "issue_stalled" in h and float(v) < 0.05: continue
            except ValueError: pass
            print(f"{h} [{u}] = {v}")
    print("-" * 60)
