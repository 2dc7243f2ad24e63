#!/bin/bash
# Round 2: parity tests that changed + A/B of the new tracer-step knobs
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_zz_series_year.py tests/test_gpu_col.py tests/test_gpu_parity.py tests/test_gpu_restart.py -m gpu -q -s --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|col vs strict\|flip rate\|worst cell" $OUT/pytest_gpu_$TAG.log | head -60
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -8 $OUT/smoke_$TAG.log
bash tools/gpu_ab_col.sh $TAG "CG_X=0" "CG_CO_SKIP=0" "CG_COL_ORDER=1" "CG_COL_ORDER=1 CG_COL_CFG=1"
