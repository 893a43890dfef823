#!/bin/bash
# round-2 session b: new S2 (short-circuit trio, needed tiles), S3 hand-over + thread-per-pair vote, kept-peak compaction
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log
for wl in ${WLS:-mini cfg4 cfg3 cfg2}; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"
  python - <<P
import json
try:
    d=json.load(open('$O/bench_$wl.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['stage_ms_per_step'], d['index_build']['gbp_per_s'], d['index_build']['device_ms'], d['result']['planted_recovered'], d['result']['peaks'], d['result']['interval_sha256'][:16], d['host_wall_ms_last_resident_step'])
except Exception as ex: print('no json', ex)
P
  tail -3 $O/bench_$wl.err
done
if [ -n "${MEMCHECK:-}" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "s3_vote_paths or s2_in_steps or marks_few or fasta_shapes" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/memcheck.log
fi
