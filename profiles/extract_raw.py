"""ncu .ncu-rep -> the small `metric,unit,value` summary committed under profiles/ (one kernel launch per report).
Usage: python profiles/extract_raw.py gpurun_out/x.ncu-rep profiles/name_raw.csv"""
import csv, subprocess, sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "smsp__inst_executed_pipe_tensor_op_dmma.sum")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for n, u, v in zip(names, units, vals):
        if n == "Kernel Name" or n in KEEP or (n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")):
            w.writerow([n, u, v])
print("wrote", sys.argv[2])
