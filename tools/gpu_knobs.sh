#!/bin/bash
# schedule / occupancy knobs re-measured under the current schedule: default bench line without the CPU arm per setting
OUT=gpurun_out; mkdir -p $OUT
run() { env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('%-40s value %.4f M ms %.3f e2e %.4f M blown %d' % ('$*', j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, j['blown_up_members']))" | tee -a $OUT/knobs_$TAG.log; }
TAG=$1
run CG_X=0
run CG_BG_SWEEP_EARLY=2
run CG_BG_SWEEP_EARLY=0
run CG_BG_SURF_MINB=3
run CG_BG_SWEEP_MINB=3
run CG_X=0
