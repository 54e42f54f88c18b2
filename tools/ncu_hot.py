#!/usr/bin/env python3
"""Print the hottest straight-line region of a kernel from an .ncu-rep (SASS + stall samples per instruction).
usage: ncu_hot.py report.ncu-rep [min_executed_fraction]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
end = his[1] - 1 if len(his) > 1 else len(rows)
body = [r for r in rows[hi + 1:end] if len(r) == len(h)]
def f(x):
    try: return int(float(x))
    except: return 0
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
ex = [f(r[ix['Instructions Executed']]) for r in body]
sm = [f(r[ix['# Samples']]) for r in body]
print('total samples', sum(sm), 'warp instructions', sum(ex))
c = collections.Counter(e for e in ex if e > 0)
print('most common execution counts', c.most_common(5))
hot = c.most_common(1)[0][0]
idx = [i for i, e in enumerate(ex) if e == hot]
print('hot count', hot, 'lines', len(idx), 'samples in hot lines', sum(sm[i] for i in idx))
start = idx[len(idx) // 2]
while 'BRA.DIV' not in body[start][ix['Source']] and start > 0: start -= 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 70
tot = 0
for r in body[start:start + n]:
    s = f(r[ix['# Samples']]); tot += s
    st = {kk[6:]: f(r[ix[kk]]) for kk in stall_cols if f(r[ix[kk]]) > 0 and kk != 'stall_selected'}
    print(f"{f(r[ix['Instructions Executed']]):7d} {s:5d} {r[ix['Source']][:56]:56s} {st}")
print('window samples', tot)
# other hot spots
print('--- other lines with many samples')
for i, r in enumerate(body):
    if sm[i] >= 0.01 * sum(sm) and ex[i] != hot:
        print(i, ex[i], sm[i], r[ix['Source']][:60])
