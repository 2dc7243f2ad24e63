#!/bin/bash
# A/B of the mix-on-write tracer step (k_ts_pre + k_tstep_col<PV>) against flux kernel + k_co_col.
#   gpurun --timeout 1200 -- 'bash tools/gpu_mix.sh r1y'
TAG=${1:-mix}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_col.py tests/test_gpu_parity.py -x -q > $OUT/pytest_mix_$TAG.log 2>&1
tail -5 $OUT/pytest_mix_$TAG.log
{
  CG_COL_MIX=0 timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile
  for mb in 2 3 4; do
    CG_COL_MIX=1 CG_PRE_MINB=$mb timeout 300 python tools/prof_run.py --members 128 --spin 400 --steps 48 --variant col --profile
  done
} > $OUT/prof_mix_$TAG.log 2>&1
cat $OUT/prof_mix_$TAG.log
for mix in 0 1; do
  CG_COL_MIX=$mix timeout 600 python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mix=$mix ms/yr %.2f e2e %.0f frac %.3f launch_ms %.4f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms']), {k: round(v,1) for k,v in d['roofline']['family_ms_per_year'].items()})" | tee -a $OUT/bench_mix_$TAG.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_col|k_ts_pre" -s 40 -c 4 \
    -o $OUT/prof_mixk_$TAG -f python tools/prof_run.py --members 128 --spin 400 --steps 4 --variant col > $OUT/prof_mixk_$TAG.log 2>&1
tail -3 $OUT/prof_mixk_$TAG.log
