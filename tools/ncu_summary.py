"""Summarise an .ncu-rep (raw page) into the handful of numbers we track: python tools/ncu_summary.py rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum"]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")
                         and float(r[i].replace(",", "") or 0) > 0.1):
            print(f"{h:85s} {units[i]:12s} {r[i]}")
