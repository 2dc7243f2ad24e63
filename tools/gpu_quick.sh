#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col.py tests/test_gpu_biogem.py -x -q 2>&1 | tail -3
run() { echo "== $1"; env $1 timeout 600 python bench.py --steps 6 --warmup 12 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/yr %.2f e2e %.0f' % (d['ms_per_step'], d['e2e']['value']), {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})"; }
run "CG_X=1" | tee -a $OUT/quick_$TAG.log
