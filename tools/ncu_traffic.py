#!/usr/bin/env python
"""Record the ncu-measured DRAM traffic of a kernel in profiles/traffic.json.

  tools/ncu_traffic.py <capture.ncu-rep> <key> [kernel-name-regex]

Reads dram__bytes_read.sum + dram__bytes_write.sum of the (first matching)
launch in an `ncu --set full` capture and stores it under <key> together with
the capture's file name and the sha16 of the kernel sources (bench.source_hash):
bench.py reports the number and flags it `stale` when the sources have changed
since the capture."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, key = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    dur = r[ix["gpu__time_duration.sum"]] + " " + units[ix["gpu__time_duration.sum"]]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[key] = {"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "duration_under_ncu": dur,
               "capture": os.path.basename(rep), "source_sha16": bench.source_hash(key)}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print(key, db[key])
    break
else:
    sys.exit("no matching launch in " + rep)
