"""Summarise an `ncu --page raw --csv` dump: python scripts/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_wait.ratio', 'smsp__average_warp_latency_issue_stalled_membar.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
for w in want:
    idx = [i for i, h in enumerate(hdr) if h == w]
    if not idx:
        print("MISSING", w); continue
    i = idx[0]
    print(f"{w:72s} {units[i]:10s}", [d[i][:48] for d in data])
