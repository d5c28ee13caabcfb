# usage: python scratch/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/rXX_ncu_*.txt
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
           "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
rep = sys.argv[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
print(f"# source report: {rep} (not committed; gpurun_out/ is scratch)")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("Kernel Name =", r[col["Kernel Name"]])
    for m in METRICS:
        if m in col:
            print(f"{m} = {r[col[m]]} {units[col[m]]}")
    print("---")
