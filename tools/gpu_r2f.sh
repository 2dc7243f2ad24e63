#!/bin/bash
# Round 2: copy issue spread over three warps, region-wise passive averaging; new bench modes (configs 1, 2, 5)
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_col.py tests/test_gpu_c_driver.py -m gpu -q -s --durations=5 -k "not century" > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
bash tools/gpu_ab_col.sh $TAG "CG_X=0" "CG_CO_PAIR=1" "CG_COL_V=2"
for C in 1 2 5; do
  timeout 600 python bench.py --config $C --steps 5 --warmup 3 > $OUT/bench_config${C}_$TAG.json 2> $OUT/bench_config${C}_$TAG.err
  python - <<EOF
import json
try:
    d = json.load(open("$OUT/bench_config${C}_$TAG.json")); r = d["roofline"]
    print("config $C: value %.4g %s, ms/step %.3f, e2e %.4g, tstepo %.1f us, %.1f GB/s (frac %.3f)" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], 1e3 * r["avg_launch_ms"], r["achieved"], r["frac"]))
except Exception as ex:
    print("config $C failed:", ex); print(open("$OUT/bench_config${C}_$TAG.err").read()[-1500:])
EOF
done
