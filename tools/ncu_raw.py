"""Prints selected raw metrics of every launch in an .ncu-rep:  python tools/ncu_raw.py file.ncu-rep"""
import csv
import subprocess
import sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
names = [r[hdr.index("Kernel Name")][:60] for r in rows[2:]]
print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(names))) + " |")
print("|---|" + "---|" * len(names))
print("| kernel | " + " | ".join(names) + " |")
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"| {k} ({rows[1][i]}) | " + " | ".join(r[i] for r in rows[2:]) + " |")
