"""Key counters of an .ncu-rep (raw page) as one text block per kernel: python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_sample_buffer_size']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    print(d.get('Kernel Name', '?'))
    for k in KEYS:
        if k in d and d[k] != '':
            print(f'  {k} = {d[k]} {u.get(k, "")}')
