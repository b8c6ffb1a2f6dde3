#!/usr/bin/env python
"""condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line
per kernel launch with the metrics DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.csv"""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('sm__inst_executed.avg.per_cycle_active', 'ipc'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
    ('sm__pipe_tensor_subunit_op_umma_cycles_active.avg.pct_of_peak_sustained_active',
     'tensor_umma_pct'),
    ('sm__inst_executed_pipe_tensor_subunit_op_umma.avg.pct_of_peak_sustained_active',
     'tensor_inst_pct'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'lanes_per_inst'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
     'stall_short_sb'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    tens = [n for n in hdr if 'tensor' in n and 'pct_of_peak_sustained_active' in n]
    keys = [k for k in KEYS if k[0] in col]
    for n in tens:
        if n not in [k[0] for k in keys]:
            keys.append((n, n.split('.')[0].replace('sm__', '')))
    w.writerow(['kernel', 'grid', 'block'] + ['%s[%s]' % (s, units[col[k]]) for k, s in keys])
    for r in rows[2:]:
        w.writerow([r[col['Kernel Name']][:60], r[col['Grid Size']], r[col['Block Size']]] +
                   [r[col[k]] for k, _ in keys])


if __name__ == '__main__':
    main(sys.argv[1])
