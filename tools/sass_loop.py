#!/usr/bin/env python
"""Static instruction mix of the loops of one kernel (no GPU needed).

  tools/sass_loop.py extensisq_b200/csrc/build/inst_Ts5.o 'rk_persistent.*Ts5.*Lorenz63' [min_len]

Disassembles with cuobjdump, finds every backward branch (a loop), and prints
for each loop of at least min_len instructions: length, the number of
fp64-pipe instructions (DFMA/DMUL/DADD/DSETP), and the opcode histogram."""
import re, subprocess, sys
from collections import Counter

obj, pat = sys.argv[1], re.compile(sys.argv[2])
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, funcs = None, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        funcs[name] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name:
        funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
for fn, ins in funcs.items():
    if not pat.search(fn):
        continue
    print("==", fn[:110], len(ins), "instructions")
    addr_ix = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t <= a and t in addr_ix:
                loops.append((addr_ix[t], i))
    for lo, hi in loops:
        if hi - lo + 1 < min_len:
            continue
        ops = Counter()
        for a, s in ins[lo:hi + 1]:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
            ops[m.group(2) if m else "?"] += 1
        n = hi - lo + 1
        nf = sum(ops[o] for o in FP64)
        print("loop %#x..%#x: %d instr, %d fp64-pipe (%s), %d other" % (
            ins[lo][0], ins[hi][0], n, nf,
            " ".join("%s %d" % (o, ops[o]) for o in FP64), n - nf))
        print("   ", " ".join("%s:%d" % kv for kv in ops.most_common(40)))
