#!/bin/bash
# Round 2 checkpoint: whole GPU suite, smoke, default bench line (512 members), e2e with two member groups, reference arm
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -14 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json')); r=d['roofline']
print('bench: value %.3fM ms/yr %.2f e2e %.3fM (h2d %.0f MB) tstepo %.1f us frac %.3f cpu %.1fk' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['h2d_bytes_per_step']/1e6, 1e3*r['avg_launch_ms'], r['frac'], d['cpu_baseline']['value']/1e3))
for k,v in r['other_families'].items(): print('   %-9s %.3f ms/call  %.0f GB/s  frac %.3f' % (k, v['avg_call_ms'], v['achieved'], v['frac']))
PY
timeout 900 python bench.py --e2e-groups 2 --steps 5 --warmup 2 --no-cpu-baseline > $OUT/bench_g2_$TAG.json 2> $OUT/bench_g2_$TAG.err
python -c "
import json
d=json.load(open('$OUT/bench_g2_$TAG.json')); print('e2e with 2 groups of 256: %.3fM (value %.3fM)' % (d['e2e']['value']/1e6, d['value']/1e6))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "
import json
d=json.load(open('$OUT/bench_ref_$TAG.json')); print('reference arm: %.1fk model-years/hour on %d cores' % (d['value']/1e3, d['cpu_baseline']['cores']))"
