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
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keep:
            print(f"{h}\t{vals[i]}\t{units[i]}")
    print()
