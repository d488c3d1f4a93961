"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > both.csv
    python profiles/ncu_lines.py both.csv [top_n]
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname, hdr = None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])   # (file, line) -> [inst, thread_inst, samples, source]
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ki, kt, ks = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]))
            agg[cur][3] = r[1].strip()[:100]
        if cur is None or len(r) <= kt or r[2] == "":
            continue
        try:
            agg[cur][0] += int(r[ki]); agg[cur][1] += int(r[kt]); agg[cur][2] += int(r[ks])
        except ValueError:
            pass
    total = sum(v[0] for v in agg.values())
    samples = sum(v[2] for v in agg.values())
    print(f"total warp instructions {total}  samples {samples}")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:>11} {100 * v[0] / total:5.1f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  smp {100 * v[2] / max(samples, 1):5.1f}%  {f}:{l}  {v[3]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
