#!/bin/bash
# round-2 session c (2+ GPUs): the multi-GPU plan against the oracle, then the strong-scaling bench
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
N=${N:-2}
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu_multi.txt 2>&1
timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q --durations=5 > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log; tail -25 $O/pytest_multi.log
for wl in ${WLS:-cfg2 cfg4}; do
  timeout 900 python bench.py --gpus $N --workload $wl --steps 3 --warmup 1 --no-cpu > $O/bench_${wl}_${N}gpu.json 2> $O/bench_${wl}_${N}gpu.err; echo "bench $wl x$N rc=$?"
  python - <<P
import json
try:
    d=json.load(open('$O/bench_${wl}_${N}gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'], d['result']['interval_sha256'][:16], d['parallelism'], d['host_wall_ms_last_resident_step'])
except Exception as ex: print('no json', ex)
P
  tail -5 $O/bench_${wl}_${N}gpu.err
done
