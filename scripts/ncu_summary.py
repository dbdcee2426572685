"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles cite."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_membar',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_drain']
import re
extra = re.compile(r"^(l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(atom|ld|st))?\.sum(\.pct_of_peak_sustained_elapsed)?|"
                   r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_atom)?\.sum|"
                   r"smsp__inst_executed_op_shared_atom[a-z_]*\.sum|smsp__inst_executed_op_global_(red|atom)\.sum|"
                   r"lts__t_sectors_op_(atom|red)\.sum|launch__shared_mem_per_block_(static|dynamic)|"
                   r"launch__occupancy_limit_[a-z_]+|launch__grid_size|launch__block_size|"
                   r"sm__inst_executed_pipe_fp64\.sum|smsp__inst_executed_pipe_fp64[a-z_]*\.sum|"
                   r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
                   r"sm__inst_executed_pipe_tensor[a-z_0-9]*\.sum)$")
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keep or extra.search(h):
            print(f"{h}\t{vals[i]}\t{units[i]}")
    print()
