"""Key counters of every kernel in an `ncu --page raw --csv` dump:  python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    print('----', d[idx['Kernel Name']][:100], 'grid', d[idx.get('Grid Size', 0)], 'block', d[idx.get('Block Size', 0)])
    for w in want:
        if w in idx:
            print(f"   {w:66s} {d[idx[w]]:>18s} {units[idx[w]]}")
