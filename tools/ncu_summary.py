"""Summarise an `ncu --page raw --csv` export: one block of key metrics per profiled launch."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
STALL = 'smsp__average_warps_issue_stalled_'

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('----')
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('%-70s %s %s' % (w, r[i][:80], units[i]))
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
            try:
                stalls.append((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    print('stalls per issued instruction:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(stalls, reverse=True)[:9]))
