#!/bin/bash
# Round 2: exact reciprocal divisions (EMBM implicit iterations, carbonate solve); parity + bench at 512 members
TAG=${1:-r2o}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_biogem.py tests/test_gpu_z_slice.py -m gpu -q -s --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
for M in ${MEMBERS:-512}; do
  timeout 900 python bench.py --members $M --steps 5 --warmup 2 --spinup-years ${SPIN:-100} --no-cpu-baseline > $OUT/bench_M${M}_$TAG.json 2> $OUT/bench_M${M}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_M${M}_$TAG.json")); r = d["roofline"]; M = $M
    print("M=%d: %.3f M my/h  ms/yr %.2f  per-member-us/yr %.1f  e2e %.3fM  tstepo %.1f us frac %.3f" % (M, d["value"]/1e6, d["ms_per_step"], 1e3*d["ms_per_step"]/M, d["e2e"]["value"]/1e6, 1e3*r["avg_launch_ms"], r["frac"]))
    print("   family us per member-year:", {k: round(1e3*v/M, 2) for k, v in r["family_ms_per_year"].items()})
except Exception as ex:
    print("M=$M failed:", ex); print(open("$OUT/bench_M${M}_$TAG.err").read()[-1500:])
PY
done
