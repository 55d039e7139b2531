"""Summarise an .ncu-rep: key throughput metrics, stall reasons, and per-phase sample distribution.
Usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import subprocess
import sys
import collections

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else '.'


def ncu(*args):
    return subprocess.run(['ncu', '-i', rep, '--kernel-name', 'regex:' + kre] + list(args),
                          capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu('--page', 'raw', '--csv'))))
hdr, units = raw[0], raw[1]
WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'gcc__cache_requests_type_instruction.sum']
for r in raw[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '')[:80])
    for k in WANT:
        if k in d:
            print('  %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = [(float(v), k) for k, v in d.items() if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio') and v]
    print('  stalls (cycles per issued instruction):',
          ', '.join('%s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
                    for v, k in sorted(st, reverse=True)[:8]))
