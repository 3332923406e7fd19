#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export to the metrics the profiles/ summaries quote (one block per kernel)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('====', r[idx['Kernel Name']][:90])
        for w in WANT:
            if w in idx:
                print('  %-75s %s %s' % (w, r[idx[w]], units[idx[w]]))
        st = [(float(r[i].replace(',', '')), h[len(STALL):-len('_per_issue_active.ratio')]) for h, i in idx.items()
              if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
        st.sort(reverse=True)
        print('  stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
