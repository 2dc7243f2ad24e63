#!/bin/bash
# Round 2: pipelined column kernel + block-cooperative convection kernel: parity tests, then A/B on the bench state
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_col.py tests/test_gpu_c_driver.py tests/test_gpu_zz_series_year.py tests/test_gpu_restart.py -m gpu -q -s --durations=5 -k "not century" > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error\|col vs strict\|flip rate\|worst cell" $OUT/pytest_gpu_$TAG.log | head -60
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -4 $OUT/smoke_$TAG.log
bash tools/gpu_ab_col.sh $TAG "CG_COL_V=1 CG_CO_V=1" "CG_COL_V=2 CG_CO_V=1" "CG_COL_V=1 CG_CO_V=2" "CG_COL_V=2 CG_CO_V=2" "CG_COL_V=2 CG_CO_V=2 CG_COL_CFG=1"
