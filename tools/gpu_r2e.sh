#!/bin/bash
# Round 2: BIOGEM parity after the libm / ordered-sum changes, bench A/B, ncu --set full of both flux / convection kernel forms
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_biogem.py tests/test_gpu_z_sig.py tests/test_gpu_c_driver.py tests/test_gpu_parity.py -m gpu -q -s --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
bash tools/gpu_ab_col.sh $TAG "CG_COL_V=1 CG_CO_V=1"
for V in 1 2; do
  CG_COL_V=$V CG_CO_V=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_col|k_co_col|k_co_blk" -s 19200 -c 4 \
    -o $OUT/prof_col_v${V}_$TAG -f python tools/prof_run.py --members 128 --spin 9600 --steps 4 --variant col > $OUT/prof_col_v${V}_$TAG.log 2>&1
  ncu -i $OUT/prof_col_v${V}_$TAG.ncu-rep --page raw --csv > $OUT/raw_col_v${V}_$TAG.csv 2>/dev/null
  ncu -i $OUT/prof_col_v${V}_$TAG.ncu-rep --page details --csv > $OUT/details_col_v${V}_$TAG.csv 2>/dev/null
  ncu -i $OUT/prof_col_v${V}_$TAG.ncu-rep --page source --csv > $OUT/source_col_v${V}_$TAG.csv 2>/dev/null
done
ls -la $OUT | tail -12
