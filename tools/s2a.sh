#!/bin/bash
# round-2 session a: existing GPU tests against the device FASTA ingest, new bench on mini / cfg4 / cfg3 with the round-1 kernels (baseline)
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt; df -h /tmp >> $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
for wl in mini cfg4 cfg3; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"
  python - <<P
import json
try:
    d=json.load(open('$O/bench_$wl.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'], d['index_build']['gbp_per_s'], d['index_build']['device_ms'], d['setup_seconds'], d['result']['planted_recovered'], d['result']['peaks'])
except Exception as ex: print('no json', ex)
P
  tail -3 $O/bench_$wl.err
done
nvidia-smi --query-gpu=memory.used --format=csv
