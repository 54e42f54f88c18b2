#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): totals per kernel, and the launches of the
last complete V-cycle.  usage: launch_summary.py launches.csv [out.txt]"""
import csv, re, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows)); hdr = r[0]; r = r[1:]
ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
def short(n):
    n = re.sub(r'\(.*', '', n).replace('void dgb::', '').replace('dgb::', '').replace('(int)', '')
    return n[:60]
L = [(short(x[ki]), x[gi], float(x[vi]) / 1e3) for x in r]
out = []
tot = collections.defaultdict(lambda: [0, 0.0])
for k, g, t in L:
    tot[k][0] += 1; tot[k][1] += t
T = sum(v[1] for v in tot.values())
out.append(f"# {len(L)} launches, {T/1e3:.1f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    out.append(f"{t/1e3:10.3f} ms {100*t/T:5.1f}%  n={n:5d} avg {t/n:9.1f} us  {k}")
# last complete cycle: between the last two "entry" helpers on the fine level that start a pre-smoother
idx = [i for i, (k, g, t) in enumerate(L) if k.startswith('k_gs_helper<9, true') or k.startswith('k_gs_helper<9, 1')]
if len(idx) >= 7:
    a, b = idx[4], idx[6]        # the third V-cycle of the run (two entry-residual helpers per cycle)
    cyc = collections.defaultdict(lambda: [0, 0.0])
    for k, g, t in L[a:b]:
        cyc[k][0] += 1; cyc[k][1] += t
    Tc = sum(v[1] for v in cyc.values())
    out.append(f"\n# one steady V-cycle: {b - a} launches, {Tc/1e3:.2f} ms of kernel time")
    for k, (n, t) in sorted(cyc.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t/1e3:10.3f} ms {100*t/Tc:5.1f}%  n={n:4d} avg {t/n:9.1f} us  {k}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
