#!/bin/bash
# A/B of the tracer-step knobs on the bench state.  gpurun --timeout 900 -- 'bash tools/gpu_ab_col.sh tag "ENV1" "ENV2" ...'
TAG=${1:-ab}; shift
OUT=gpurun_out
mkdir -p $OUT
run() { echo "== $1"; env $1 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>>$OUT/ab_$TAG.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/yr %.2f  e2e %.3fM  tstepo %.1f us  frac %.3f ' % (d['ms_per_step'], d['e2e']['value']/1e6, 1e3*r['avg_launch_ms'], r['frac']), {k: round(v,1) for k,v in r['family_ms_per_year'].items()})"; }
for envs in "$@"; do run "$envs"; done 2>&1 | tee $OUT/ab_$TAG.log
