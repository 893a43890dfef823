#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep (source page): stall mix, top instructions, instruction-count blocks."""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isamp, iex, ithr = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in body); totex = sum(int(r[iex]) for r in body)
print('total samples', tot, 'total warp instr', totex)
agg = {}
for r in body:
    for i in cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot * 0.01})
for r in sorted(body, key=lambda r: -int(r[isamp]))[:ntop]:
    print(r[0][-5:], r[isamp], f"{100 * int(r[isamp]) / tot:5.1f}%", r[iex], r[ithr], r[1][:70], {hdr[i]: r[i] for i in cols if int(r[i] or 0) > int(r[isamp]) * 0.3})
groups = []
for r in body:
    ex = int(r[iex])
    if groups and groups[-1][0] == ex: groups[-1][1] += 1; groups[-1][3] += int(r[isamp])
    else: groups.append([ex, 1, r[0][-5:], int(r[isamp]), r[1][:50]])
t = sum(g[0] * g[1] for g in groups)
for g in groups:
    if g[0] * g[1] > t * 0.015: print(f"{g[2]} exec/inst={g[0]:>11} n_inst={g[1]:4d} warp-instr={g[0]*g[1]/1e6:9.1f}M ({100*g[0]*g[1]/t:4.1f}%) samples={g[3]:7d} first={g[4]}")
