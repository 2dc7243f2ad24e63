#!/bin/bash
# Kernel experiment session: col-variant parity tests, per-family CUDA-event timings of the tracer variants / launch
# shapes on a ~4-year-old 128-member state, one ncu --set full capture of the fused column kernel.
#   gpurun --timeout 1200 -- 'bash tools/gpu_exp.sh r1g'
TAG=${1:-exp}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_col.py -x -q > $OUT/pytest_col_$TAG.log 2>&1
tail -15 $OUT/pytest_col_$TAG.log
{
  timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant fast --profile
  for c in 0 1; do
    CG_COL_CFG=$c timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile
  done
  CG_COL_CFG=0 timeout 300 python tools/prof_run.py --members 64 --spin 400 --steps 48 --variant col --profile
  CG_COL_CFG=0 timeout 300 python tools/prof_run.py --members 32 --spin 400 --steps 48 --variant col --profile
} > $OUT/prof_variants_$TAG.log 2>&1
cat $OUT/prof_variants_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_col|k_co_col|k_baro_reg" -s 30 -c 6 \
    -o $OUT/prof_col_$TAG -f python tools/prof_run.py --members 128 --spin 400 --steps 6 --variant col > $OUT/prof_col_$TAG.log 2>&1
ls -la $OUT | tail -5
