"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept under
profiles/ (kernel, launches, total ms, share of the captured device time, grid, block).

    python tools/launch_summary.py gpurun_out/r1_launches_final.csv > profiles/r1_launches_final.md
"""
import collections
import csv
import sys


def summarise(path: str, top: int = 24) -> str:
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki]
        name = name[5:].split("(")[0] if name.startswith("void ") else name.split("(")[0]
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = [f"Total device time captured: {tot:.2f} ms over {len(data)} launches.\n",
           "| kernel | launches | total ms | share | grid | block |", "|---|---:|---:|---:|---|---|"]
    for k, (c, ms, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| `{k[:100]}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% | {g} | {b} |")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    sys.stdout.write(summarise(sys.argv[1]))
