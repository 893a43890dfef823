#!/bin/bash
# per-kernel times of one cfg4 step (ncu launch list, cold-cache serialised figures: shares, not absolutes)
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
WL=${WL:-cfg4}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:lhgt -c ${NK:-3000} --csv --log-file $O/launches_$WL.csv \
   python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu > $O/ncu_bench_$WL.log 2>&1; echo "ncu list rc=$?"
python - <<P
import csv, collections
rows=[r for r in csv.reader(open('$O/launches_$WL.csv')) if len(r)>5]
hdr=None
agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: v=float(d['Metric Value'].replace(',',''))
    except: continue
    unit=d.get('Metric Unit','')
    if unit.startswith('us'): v/=1000
    elif unit.startswith('ns'): v/=1e6
    elif unit.startswith('s') and not unit.startswith('ms'): v*=1000
    k=d['Kernel Name'][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v for _,v in agg.values())
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f"{v:10.2f} ms {n:6d} x  {100*v/tot:5.1f}%  {k}")
print('total', tot)
P
