#!/bin/bash
# multi-GPU bench lines: N=${N}, workloads ${WLS}
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
N=${N:-8}
for wl in ${WLS:-cfg4 cfg5}; do
  timeout 1200 python bench.py --gpus $N --workload $wl --steps ${STEPS:-3} --warmup 1 --no-cpu ${EXTRA:-} > $O/bench_${wl}_${N}gpu.json 2> $O/bench_${wl}_${N}gpu.err; echo "bench $wl x$N rc=$?"
  python - <<P
import json
try:
    txt=open('$O/bench_${wl}_${N}gpu.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'], d['result']['interval_sha256'][:16], d['parallelism'])
    print(d.get('index_sweep'))
except Exception as ex: print('no json', ex)
P
  tail -5 $O/bench_${wl}_${N}gpu.err
done
