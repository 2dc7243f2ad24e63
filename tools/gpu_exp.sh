#!/bin/bash
# Kernel experiment session: GPU parity tests, per-family CUDA-event timings of the tracer variants, short bench runs
# with and without the forked momentum branch, one ncu --set full capture of the col-variant kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_exp.sh r1g'
TAG=${1:-exp}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
tail -5 $OUT/pytest_gpu_$TAG.log
{
  timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant fast --profile
  timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile
} > $OUT/prof_variants_$TAG.log 2>&1
cat $OUT/prof_variants_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_fork_$TAG.json 2> $OUT/bench_fork_$TAG.err
CG_NOFORK=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_nofork_$TAG.json 2> $OUT/bench_nofork_$TAG.err
python - <<PY
import json
for n in ("fork", "nofork"):
    try:
        d = json.load(open("$OUT/bench_%s_$TAG.json" % n))
        print(n, "value %.0f ms_per_step %.2f e2e %.0f roofline %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]), d["roofline"]["family_ms_per_year"])
    except Exception as ex:
        print(n, "failed", ex)
PY
tail -3 $OUT/bench_fork_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_col|k_co_col|k_baro_reg|k_bg_step" -s 40 -c 8 \
    -o $OUT/prof_col_$TAG -f python tools/prof_run.py --members 128 --spin 400 --steps 6 --variant col > $OUT/prof_col_$TAG.log 2>&1
