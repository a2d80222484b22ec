"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table
(launches, total ms, ms per step, share of the step kernels).  usage: launch_summary.py launches.csv n_steps > out.md
n_steps = how many steps the profiled command ran (warm-up + timed + the eager per-kernel pass)."""
import collections
import csv
import re
import sys


def main():
    path, n_steps = sys.argv[1], int(sys.argv[2])
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iK, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    n = 0
    for r in rd:
        if len(r) <= iV:
            continue
        v = float(r[iV].replace(',', ''))
        ms = v / 1e6 if r[iU] in ('nsecond', 'ns') else (v / 1e3 if r[iU] in ('usecond', 'us') else v)
        name = re.sub(r'[<(].*', '', r[iK]).replace('void ', '').replace('han::', '')
        a = agg.setdefault(name or '(memset / memcpy)', [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    ours = {k: v for k, v in agg.items() if not k.startswith('at::') and not k.startswith('at_cuda') and 'cub' not in k and k != '(memset / memcpy)'}
    setup = {'transpose_fill_kernel', 'seg_sort_warp_kernel', 'seg_sort_block_kernel', 'col_count_kernel', 'chunk_rows_kernel',
             'scan_apply_kernel', 'scan_chunk_sums_kernel', 'scan_chunk_offsets_kernel'}
    step = {k: v for k, v in ours.items() if k not in setup}
    tot = sum(v[1] for v in step.values())
    print(f'{n} launches captured; {n_steps} steps (warm-up, timed, eager per-kernel pass).  Times under ncu are cold-cache and')
    print('serialised: the SHARES are the evidence, bench.py\'s CUDA-event times the absolute numbers.\n')
    print('| kernel (ours, inside the step) | launches | total ms | ms / step | share |')
    print('|---|---|---|---|---|')
    for k, (c, ms) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {c} | {ms:.3f} | {ms / n_steps:.3f} | {ms / tot * 100:.1f} % |')
    print(f'| **sum** | | {tot:.3f} | {tot / n_steps:.3f} | |')
    print('\n| one-off graph build (ours) | launches | total ms |')
    print('|---|---|---|')
    for k, (c, ms) in sorted(((k, v) for k, v in ours.items() if k in setup), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {c} | {ms:.3f} |')
    lib = {k: v for k, v in agg.items() if k not in ours}
    print('\n| library kernels (torch: synthetic graph generation, L2 term over the variables, copies) | launches | total ms |')
    print('|---|---|---|')
    for k, (c, ms) in sorted(lib.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f'| `{k}` | {c} | {ms:.3f} |')


if __name__ == '__main__':
    main()
