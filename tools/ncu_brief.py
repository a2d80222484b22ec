"""Prints the handful of ncu raw-page metrics that decide what bounds a kernel, and the executed-instruction
share per CUDA source line (needs -lineinfo and --import-source on).  usage: ncu_brief.py report.ncu-rep [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    sel = ['--kernel-name', 'regex:' + kre] if kre else []
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv'] + sel))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('----', r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:85s} {r[i][:24]:>24s} {units[i]}')
    src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'] + sel))))
    agg, fname, iI = collections.OrderedDict(), None, None
    for r in src:
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif len(r) > 2 and r[0] == 'Line No':
            iI = r.index('Instructions Executed')
        elif len(r) > 2 and r[2] == '-' and iI is not None:
            try:
                n = int(r[iI])
            except ValueError:
                n = 0
            key = (fname, int(r[0]))
            agg[key] = (agg.get(key, (0, ''))[0] + n, r[1])
    tot = sum(v[0] for v in agg.values()) or 1
    print('executed warp instructions by source line (top 30), total', tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f'  {k[0]}:{k[1]:4d} {v[0] / tot * 100:5.1f}%  {v[1][:100]}')


if __name__ == '__main__':
    main()
