#!/bin/bash
# generic A/B: gpu_ab.sh TAG "ENV1" "ENV2" ...   (each ENV is a space-separated list of VAR=value)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for e in "$@"; do
  echo "== $e" | tee -a $OUT/ab_$TAG.log
  env $e timeout 600 python -m pytest tests/test_gpu_col.py -x -q 2>&1 | tail -1 | tee -a $OUT/ab_$TAG.log
  env $e timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile 2>&1 | grep us/step | tee -a $OUT/ab_$TAG.log
  env $e timeout 600 python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/yr %.2f e2e %.0f frac %.3f launch_ms %.4f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']), {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})" | tee -a $OUT/ab_$TAG.log
done
