#!/bin/bash
# A/B of the split column kernel (two threads per member-column) against the one-thread-per-member form.
TAG=${1:-split}
OUT=gpurun_out
mkdir -p $OUT
CG_COL_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_col.py -x -q > $OUT/pytest_split_$TAG.log 2>&1
tail -5 $OUT/pytest_split_$TAG.log
{
  for sp in 0 1 2; do
    echo "== CG_COL_SPLIT=$sp"
    CG_COL_SPLIT=$sp timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile
  done
} > $OUT/prof_split_$TAG.log 2>&1
cat $OUT/prof_split_$TAG.log
for sp in 0 1; do
  CG_COL_SPLIT=$sp timeout 600 python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('split=$sp ms/yr %.2f e2e %.0f frac %.3f launch_ms %.4f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']), {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})" | tee -a $OUT/bench_split_$TAG.log
done
CG_COL_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_split" -s 20 -c 2 \
    -o $OUT/prof_splitk_$TAG -f python tools/prof_run.py --members 128 --spin 400 --steps 4 --variant col > $OUT/prof_splitk_$TAG.log 2>&1
tail -3 $OUT/prof_splitk_$TAG.log
