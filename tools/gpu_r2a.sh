#!/bin/bash
# Round 2, first GPU session: full GPU test suite (the series-year test un-xfailed, at both call points), the call-point
# diagnostic, the default bench line.  gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh r2a'
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/nproc_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -40 $OUT/pytest_gpu_$TAG.log
timeout 300 python tools/dbg_callpoint.py 6 col > $OUT/dbg_callpoint_$TAG.log 2>&1
cat $OUT/dbg_callpoint_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
cat $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
