"""Counts the SASS opcodes that prove which hardware paths each kernel uses (cuobjdump -sass of the built library):
UTCHMMA (tcgen05.mma), UTMALDG (TMA loads), LDGSTS (cp.async), LDTM / STTM (tcgen05.ld / st), FFMA2 / FMUL2
(packed fp32), REDG / ATOMG (global reductions / atomics), HMMA (mma.sync), SYNCS (mbarrier), MUFU.  usage: sass_opcodes.py [lib.so] > out.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'UTCMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'LDGSTS', 'FFMA2', 'FMUL2',
         'FADD2', 'HMMA', 'REDG', 'ATOMG', 'ATOMS', 'SYNCS', 'MUFU', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'FFMA']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'han_b200', 'libhan_sm100.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            op = m.group(1)
            cur['_total'] += 1
            for w in WATCH:
                if op == w or op.startswith(w + '.') or (w.startswith('UTC') and op.startswith(w)):
                    cur[w] += 1
                    break
    demangled = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(per, demangled))
    print('| kernel | SASS instructions | tensor core (UTC*MMA) | TMA (UTMALDG) | TMEM (LDTM/STTM) | cp.async (LDGSTS) | FFMA2/FMUL2 | HMMA (mma.sync) | REDG / ATOMG | MUFU | SYNCS (mbarrier) |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for k, c in per.items():
        name = re.sub(r'\(.*', '', names.get(k, k)).replace('void ', '').replace('han::', '')
        utc = sum(v for w, v in c.items() if w.startswith('UTC') and 'MMA' in w)
        if c['_total'] < 50:
            continue
        print(f"| `{name[:70]}` | {c['_total']} | {utc} | {c['UTMALDG']} | {c['LDTM'] + c['STTM']} | {c['LDGSTS']} | "
              f"{c['FFMA2'] + c['FMUL2'] + c['FADD2']} | {c['HMMA']} | {c['REDG']} / {c['ATOMG']} | {c['MUFU']} | {c['SYNCS']} |")


if __name__ == '__main__':
    main()
