#!/bin/bash
# bench cfg4/cfg2 + ncu --set full captures of the hot kernels of one cfg4 step
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
WLS="${WLS:-cfg4 cfg2}" MEMCHECK= bash tools/s2b.sh 2>&1 | grep -v "^\.\|passed\|pytest rc\|^=\|call  "
for kn in ${KERNELS:-s3_vote_kernel s2_regemit_kernel s2_regapply_kernel s2_gsemit_kernel s2_gsapply_kernel s3_pairs_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$kn -s ${SKIP:-3} -c 1 -f -o $O/$kn \
     python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu > $O/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
done
ls -la $O/*.ncu-rep
