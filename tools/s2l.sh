#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, default bench (cfg4, with cpu_baseline), reference arm
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
( time timeout 1500 python bench.py --steps ${STEPS:-10} --warmup ${WARM:-3} > $O/bench_default.json 2> $O/bench_default.err ); echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads([l for l in open('$O/bench_default.json').read().splitlines() if l.startswith('{')][-1])
    print(d['config']['workload'][:60]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pcie_gbs']); print(d['roofline']['stage_ms_per_step']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['whole_step']['frac']); print(d['cpu_baseline']); print(d['clocks']); print(d['gpu_launches'], d['index_build']['gbp_per_s'])
except Exception as ex: print('no json', ex)
P
tail -3 $O/bench_default.err
if [ -n "${MEMCHECK:-}" ]; then timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "s3_vote_paths or s2_in_steps or bucketed or image_blocks or next_sample or marks_few or many_contigs" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck.log; fi
( time timeout 1500 python bench.py --impl reference --steps ${RSTEPS:-2} --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err ); echo "ref rc=$?"; cat $O/bench_ref.json | cut -c1-2500; tail -3 $O/bench_ref.err
