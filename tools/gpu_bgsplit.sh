#!/bin/bash
TAG=${1:-bgs}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col.py::test_concurrent_schedule_bit_identical tests/test_gpu_biogem.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_bgs_$TAG.log
for e in "CG_BG_SPLIT=0" "CG_BG_SPLIT=1" "CG_BG_SPLIT=1 CG_BG_SURF_MINB=3" "CG_BG_PIPE=1 CG_BG_SPLIT=0"; do
  echo "== $e" | tee -a $OUT/ab_$TAG.log
  env $e timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/yr %.2f e2e %.0f frac %.3f launches %d' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']), {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})" | tee -a $OUT/ab_$TAG.log
done
