#!/bin/bash
# Builds variants of liblhgt.so with -D overrides (here, no GPU needed) or times them on the GPU box.
#   tools/sweep.sh build            -> tools/_build/variants/<name>.so
#   tools/sweep.sh run              -> one `bench.py --no-cpu` line per variant in gpurun_out/sweep_<name>.json
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
V=tools/_build/variants
declare -A VAR=(
  [base]=""
  [bin_w16_c2]="-DLHGT_BIN_WARPS=16 -DLHGT_BIN_CTAS=2"
  [bin_w32_c1]="-DLHGT_BIN_WARPS=32 -DLHGT_BIN_CTAS=1"
  [bin_w16_c2_s3_w16_c2]="-DLHGT_BIN_WARPS=16 -DLHGT_BIN_CTAS=2 -DLHGT_S3_WARPS=16 -DLHGT_S3_CTAS=2"
)
case "${1:-build}" in
build)
  mkdir -p $V
  for n in "${!VAR[@]}"; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -Xptxas -warn-spills ${VAR[$n]} \
      -o $V/$n.so localhgt_b200/csrc/lhgt_kernels.cu localhgt_b200/csrc/lhgt_api.cu 2>&1 | grep -E "error|ILi3.*spill|s1_leaf.*spill" ; echo "built $n"
  done;;
run)
  mkdir -p gpurun_out
  for n in "${!VAR[@]}"; do
    LHGT_LIB=$PWD/$V/$n.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/sweep_$n.json 2> gpurun_out/sweep_$n.err
    python -c "
import json;d=json.load(open('gpurun_out/sweep_$n.json'));s=d['roofline']['stage_ms_per_step'];print('$n', round(d['ms_per_step'],2), s['s1_hash_streams'], s['s1_split_streams'], s['s1_apply_leaves'], s['s3_pairs'])" || tail -3 gpurun_out/sweep_$n.err
  done;;
esac
