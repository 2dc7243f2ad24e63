#!/bin/bash
# quick A/B on the bench state: BIOGEM fusion / prefetch knobs
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $1"; env $1 timeout 600 python bench.py --steps 6 --warmup 12 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/yr %.2f' % d['ms_per_step'], {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})"; }
run "CG_X=0" | tee -a $OUT/quick_$TAG.log
run "CG_BG_NOFUSE=1" | tee -a $OUT/quick_$TAG.log
run "CG_BG_NOPF=1" | tee -a $OUT/quick_$TAG.log
run "CG_BG_NOFUSE=1 CG_BG_NOPF=1" | tee -a $OUT/quick_$TAG.log
