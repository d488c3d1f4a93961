"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/launch_summary.py profiles/r01_final_launches.csv > profiles/r01_final_launches_summary.txt
"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
        agg[r[ik]][0] += v
        agg[r[ik]][1] += 1
    total = sum(v for v, _ in agg.values())
    for name, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{v:10.1f} us {100 * v / total:5.1f}%  n={n:<4d} avg {v / n:8.1f} us  {name[:64]}")


if __name__ == "__main__":
    main(sys.argv[1])
