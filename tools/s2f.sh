#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
WLS="${WLS:-cfg4 cfg3}" MEMCHECK= bash tools/s2b.sh 2>&1 | grep -v "^\.\|passed\|pytest rc\|^=\|call  "
WL=cfg4 NK=400 bash tools/s2e.sh
