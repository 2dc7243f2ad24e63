#!/bin/bash
# Round 2: how do the kernel families scale with the members per handle?  Generic-shape ("fast") tracer kernels at 128 / 256 / 512
# members (the column kernel exists for <= 128), family times of the instrumented pass per member.
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_z_slice.py -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1
tail -5 $OUT/pytest_gpu_$TAG.log
for M in 128 256 512; do
  timeout 900 python bench.py --members $M --variant fast --steps 3 --warmup 1 --spinup-years 3 --no-cpu-baseline > $OUT/bench_fast_M${M}_$TAG.json 2> $OUT/bench_fast_M${M}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_fast_M${M}_$TAG.json")); r = d["roofline"]; M = $M
    print("M=%d: %.3f M my/h  ms/yr %.2f  per-member-us/yr %.1f  e2e %.3fM" % (M, d["value"]/1e6, d["ms_per_step"], 1e3*d["ms_per_step"]/M, d["e2e"]["value"]/1e6))
    print("   family us per member-year:", {k: round(1e3*v/M, 2) for k, v in r["family_ms_per_year"].items()})
except Exception as ex:
    print("M=$M failed:", ex); print(open("$OUT/bench_fast_M${M}_$TAG.err").read()[-1500:])
PY
done
