#!/bin/bash
# A/B of environment knobs on the default bench line:  bash tools/gpu_ab.sh TAG "A=1" "B=2 C=3" ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for SET in "$@"; do
  echo "== $SET"
  env $SET timeout 900 python bench.py --steps ${STEPS:-5} --warmup 2 --spinup-years ${SPIN:-100} --no-cpu-baseline ${BENCH_ARGS} > $OUT/bench_ab_$TAG.json 2> $OUT/bench_ab_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_ab_$TAG.json")); r = d["roofline"]; M = d["config"]["members_per_gpu"]
    print("M=%d: %.3f M my/h  ms/yr %.2f  e2e %.3fM  tstepo %.1f us frac %.3f" % (M, d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, 1e3*r["avg_launch_ms"], r["frac"]))
    print("   family us per member-year:", {k: round(1e3*v/M, 2) for k, v in r["family_ms_per_year"].items()})
except Exception as ex:
    print("failed:", ex); print(open("$OUT/bench_ab_$TAG.err").read()[-1500:])
PY
done
