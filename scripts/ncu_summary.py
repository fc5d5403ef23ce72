"""Summarise an .ncu-rep: per kernel key metrics (raw page) and, with --source KERNEL, the hottest source lines."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('--- %s  id=%s' % (r[hdr.index('Kernel Name')][:70], r[hdr.index('ID')]))
    for w in want:
        if w in hdr:
            print('   %-70s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 8:
                print('   STALL %-60s %.1f' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
