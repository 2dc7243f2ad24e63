#!/bin/bash
# Round 2: warp-tile column kernel (CG_COL_V=3): parity + A/B + ncu
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
CG_COL_V=3 timeout 600 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_col.py tests/test_gpu_wet_exchange.py tests/test_gpu_hosing.py -m gpu -q -s -k "not century and not launch_count" > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
bash tools/gpu_ab_col.sh $TAG "CG_COL_V=1" "CG_COL_V=3" "CG_COL_V=1 CG_CO_SKIP=0"
CG_COL_V=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_colw" -s 9600 -c 2 \
    -o $OUT/prof_colw_$TAG -f python tools/prof_run.py --members 128 --spin 9600 --steps 4 --variant col --perturb > $OUT/prof_colw_$TAG.log 2>&1
ncu -i $OUT/prof_colw_$TAG.ncu-rep --page details --csv > $OUT/details_colw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_colw_$TAG.ncu-rep --page raw --csv > $OUT/raw_colw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_colw_$TAG.ncu-rep --page source --csv > $OUT/source_colw_$TAG.csv 2>/dev/null
rm -f $OUT/prof_colw_$TAG.ncu-rep
