#!/usr/bin/env python
"""Dynamic instruction mix from an `ncu --page source --csv` dump."""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ithr, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
mix = defaultdict(float); samp = defaultdict(float); tot = 0; thr = 0
lines = []
for r in rows[2:]:
    try: ex = float(r[iex]); th = float(r[ithr]); sm = float(r[ism])
    except: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    op = m.group(2) if m else "?"
    mix[op] += ex; samp[op] += sm; tot += ex; thr += th
    lines.append((ex, sm, th / ex if ex else 0, r[isrc].strip()))
print("total warp-inst %.3e  avg active threads %.2f" % (tot, thr / tot))
ssum = sum(samp.values())
for op, v in sorted(mix.items(), key=lambda x: -x[1])[:28]:
    print("%-10s %6.2f%% inst   %6.2f%% samples" % (op, 100 * v / tot, 100 * samp[op] / ssum))
if len(sys.argv) > 2:
    print("---- hottest by samples")
    for ex, sm, at, src in sorted(lines, key=lambda x: -x[1])[:int(sys.argv[2])]:
        print("%10.3e ex %6d smp  thr %4.1f  %s" % (ex, sm, at, src[:90]))
