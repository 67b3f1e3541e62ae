#!/usr/bin/env python
"""Per-source-line executed warp instructions from `ncu --page source --csv
--print-source cuda,sass`; with two files prints the lines that differ most."""
import csv, sys


def load(path):
    rows = list(csv.reader(open(path)))
    cur, agg = None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0].isdigit():
            try:
                ex, sm = float(r[7]), float(r[6])
            except ValueError:
                continue
            k = (cur, int(r[0]))
            a = agg.get(k, (0.0, 0.0, r[1]))
            agg[k] = (a[0] + ex, a[1] + sm, r[1])
    return agg


a = load(sys.argv[1])
ta = sum(v[0] for v in a.values())
if len(sys.argv) > 2 and not sys.argv[2].isdigit():
    b = load(sys.argv[2])
    tb = sum(v[0] for v in b.values())
    print("total %.4e vs %.4e  (delta %.3e)" % (ta, tb, ta - tb))
    keys = set(a) | set(b)
    d = sorted(((a.get(k, (0, 0, ""))[0] - b.get(k, (0, 0, ""))[0], k) for k in keys),
               key=lambda x: -abs(x[0]))
    for dv, k in d[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        src = (a.get(k) or b.get(k))[2]
        print("%+10.3e  %s:%d  %s" % (dv, k[0], k[1], src.strip()[:80]))
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print("total %.4e" % ta)
    for k, v in sorted(a.items(), key=lambda x: -x[1][0])[:n]:
        print("%10.3e %5.2f%% smp %6d  %s:%d  %s" % (v[0], 100 * v[0] / ta, v[1], k[0], k[1], v[2].strip()[:70]))
