#!/usr/bin/env python3
"""Instruction counts per kernel from `cuobjdump -sass` of libdgb200.so (evidence for TMA bulk copies, mbarriers,
cluster instructions, FP64 tensor-core use).  usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "dg_multigrid_solver_b200", "libdgb200.so")
OPS = ["UBLKCP", "SYNCS", "MAPA", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "DMMA", "UTMALDG",
       "ATOM", "BAR", "ST.E", "LD.E", "CCTL", "MEMBAR", "ERRBAR", "NANOSLEEP", "UCGABAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        total[cur] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o in ("ST.E", "LD.E") and op.startswith(o)):
                if o in ("ST.E", "LD.E") and (op.startswith("STG") or op.startswith("LDG")):
                    continue
                counts[cur][o] += 1
dem = subprocess.run(["cu++filt"] + list(total), capture_output=True, text=True).stdout.splitlines()
for mangled, d in zip(list(total), dem):
    d = d.replace("(int)", "").replace("(bool)", "").replace("void dgb::", "").replace("dgb::", "")
    d = re.sub(r"\(.*", "", d)
    names[mangled] = d
print("# SASS summary of libdgb200.so (cuobjdump -sass, sm_100a cubins; tools/sass_summary.py): instruction counts per kernel")
print("# kernel | instructions | " + " ".join(OPS))
for mangled in sorted(total, key=lambda k: names[k]):
    c = counts[mangled]
    print(f"{names[mangled][:70]:70s} {total[mangled]:6d}  " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]))
