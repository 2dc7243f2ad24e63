#!/bin/bash
# quick A/B: col-variant tests + per-family timings (no ncu)
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_col.py tests/test_gpu_biogem.py -x -q 2>&1 | tail -3
timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile 2>&1 | tee $OUT/prof_quick_$TAG.log
