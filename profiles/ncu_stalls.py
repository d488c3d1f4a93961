"""Per-source-line instruction counts and stall samples from an
`ncu --page source --csv --print-source cuda,sass` export (SASS rows only, no double counting).

    python profiles/ncu_stalls.py both.csv <n_env_steps_per_launch> [top_n]
"""
import csv
import sys
from collections import defaultdict


def main(path, units, top=30):
    rows = list(csv.reader(open(path)))
    fname = hdr = cur = None
    agg, smp, src = defaultdict(int), defaultdict(int), {}
    stall = defaultdict(lambda: defaultdict(int))
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
            ki, ks = hdr.index("Instructions Executed"), hdr.index("# Samples")
            sidx = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]))
            src[cur] = r[1].strip()[:90]
            continue
        try:
            agg[cur] += int(r[ki])
            smp[cur] += int(r[ks])
            for i, h in sidx:
                if r[i] not in ("", "0"):
                    stall[cur][h] += int(r[i])
        except ValueError:
            pass
    tot, ts = sum(agg.values()), sum(smp.values())
    print(f"warp instructions {tot}  = {tot / units:.1f} per unit; samples {ts}")
    byfile = defaultdict(int)
    for k, v in agg.items():
        byfile[k[0]] += v
    print({f: round(v / units, 1) for f, v in byfile.items()})
    allst = defaultdict(int)
    for d in stall.values():
        for h, v in d.items():
            allst[h] += v
    print({h: round(100 * v / ts, 1) for h, v in sorted(allst.items(), key=lambda kv: -kv[1])})
    print("--- by instructions")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"inst {v / units:6.1f}/u {100 * v / tot:5.1f}%  smp {100 * smp[k] / ts:5.1f}%  {k[0]}:{k[1]}  {src[k]}")
    print("--- by stall samples")
    for k, v in sorted(smp.items(), key=lambda kv: -kv[1])[:top]:
        t2 = sorted(stall[k].items(), key=lambda kv: -kv[1])[:2]
        print(f"smp {100 * v / ts:5.1f}%  inst {agg[k] / units:6.1f}/u  {k[0]}:{k[1]}  {t2}  {src[k][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 30)
