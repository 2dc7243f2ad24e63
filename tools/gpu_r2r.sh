#!/bin/bash
# Round 2: EMBM forms (both fields side by side; 3 cells x 448 threads against 2 cells x 672 threads), fused coupling at 512 members
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head
CG_EMBM_CPT2=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k embm >> $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | tail -2
run() {
  echo "== $*"
  env "$@" timeout 900 python bench.py --steps 5 --warmup 2 --spinup-years ${SPIN:-100} --no-cpu-baseline > $OUT/bench_ab_$TAG.json 2> $OUT/bench_ab_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_ab_$TAG.json")); r = d["roofline"]; M = d["config"]["members_per_gpu"]
    print("M=%d: %.3f M my/h  ms/yr %.2f  e2e %.3fM  tstepo %.1f us frac %.3f" % (M, d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, 1e3*r["avg_launch_ms"], r["frac"]))
    print("   family us per member-year:", {k: round(1e3*v/M, 2) for k, v in r["family_ms_per_year"].items()})
except Exception as ex:
    print("failed:", ex); print(open("$OUT/bench_ab_$TAG.err").read()[-1500:])
PY
}
run CG_X=0
run CG_EMBM_CPT2=1
run CG_BENCH_FUSE=1
