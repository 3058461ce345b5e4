"""Pick the metrics that matter out of `ncu --page raw --csv` (stdin)."""
import csv
import sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
units = rows[1] if len(rows) > 2 else [""] * len(hdr)
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__waves_per_multiprocessor', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_xu.sum']
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print('  %-66s %s %s' % (h, vals[i], units[i]))
