#!/bin/bash
# Round 2: GPU test suite + smoke + call-point diagnostic.  gpurun --timeout 1500 -- 'bash tools/gpu_r2b.sh r2b'
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1300 python -m pytest tests -m gpu -q --durations=8 -s > $OUT/pytest_gpu_$TAG.log 2>&1
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -8 $OUT/smoke_$TAG.log
timeout 300 python tools/dbg_callpoint.py 4 col > $OUT/dbg_callpoint_$TAG.log 2>&1
cat $OUT/dbg_callpoint_$TAG.log
