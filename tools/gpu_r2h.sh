#!/bin/bash
# Round 2: convection kernel with shared-memory decision arrays; two-group end-to-end leg; ncu of the perturbed state
TAG=${1:-r2h}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_col.py -m gpu -q -s -k "not century" > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
bash tools/gpu_ab_col.sh $TAG "CG_X=0" "CG_CO_LOCAL=1" "CG_CO_PAIR=1"
env timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --e2e-groups 1 2>>$OUT/ab_$TAG.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e-groups 1: ms/yr %.2f e2e %.3fM' % (d['ms_per_step'], d['e2e']['value']/1e6))" | tee -a $OUT/ab_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_co_col" -s 9600 -c 2 \
    -o $OUT/prof_co_$TAG -f python tools/prof_run.py --members 128 --spin 9600 --steps 4 --variant col --perturb > $OUT/prof_co_$TAG.log 2>&1
ncu -i $OUT/prof_co_$TAG.ncu-rep --page details --csv > $OUT/details_co_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_co_$TAG.ncu-rep --page source --csv > $OUT/source_co_$TAG.csv 2>/dev/null
rm -f $OUT/prof_co_$TAG.ncu-rep
