#!/bin/bash
# Short end-of-round session on one B200 (no ncu: the kernels are those of the last full session): GPU tests, the default
# bench line, the reference arm, the schedule trace.   gpurun --timeout 900 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/nproc_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_ref_$TAG.json
timeout 300 python tools/trace_run.py > $OUT/trace_$TAG.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
