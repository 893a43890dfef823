#!/bin/bash
# One GPU-box session: parity tests, bench lines, ncu launch list + full captures.  Outputs -> gpurun_out/.
# usage: tools/gpu_round.sh [tests] [bench] [ncu] [full]
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
what="${*:-tests bench ncu full}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
for w in $what; do case $w in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log;;
quick)
  timeout 900 python -m pytest tests -m gpu -x -q -k "${TESTK:-s1 or streamed or k32 or large_run_properties}" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?" >> $O/pytest_quick.log; tail -5 $O/pytest_quick.log;;
benchq)
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ${BENCHARGS:-} > $O/bench_quick${TAG:-}.json 2> $O/bench_quick${TAG:-}.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$O/bench_quick${TAG:-}.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['stage_ms_per_step'])"; tail -5 $O/bench_quick${TAG:-}.err;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log;;
bench)
  timeout 1500 python bench.py --steps 5 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; tail -c 3000 $O/bench_cfg2.json; tail -5 $O/bench_cfg2.err;;
refarm)
  timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cat $O/bench_ref.json;;
ncu)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
     python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?";;
full)
  for kn in ${KERNELS:-s1_bin_kernel s1_split_kernel s1_leaf_kernel s3_pairs_kernel}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 1 -c 1 -f -o $O/$kn \
       python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
  done;;
esac; done
ls -la $O
