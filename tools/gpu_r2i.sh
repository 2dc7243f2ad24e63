#!/bin/bash
# Round 2: whole GPU suite, smoke, default bench line + reference arm, compute-sanitizer over smoke()
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
timeout 1300 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -3 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python -c "
import json
d=json.load(open('$OUT/bench_$TAG.json')); r=d['roofline']
print('bench: value %.3fM ms/yr %.2f e2e %.3fM (h2d %.0f MB) tstepo %.1f us frac %.3f cpu %.1fk' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['h2d_bytes_per_step']/1e6, 1e3*r['avg_launch_ms'], r['frac'], d['cpu_baseline']['value']/1e3))
for k,v in r['other_families'].items(): print('   %-9s %.3f ms/call  %.0f GB/s  frac %.3f' % (k, v['avg_call_ms'], v['achieved'], v['frac']))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "
import json
d=json.load(open('$OUT/bench_ref_$TAG.json')); print('reference arm: %.1fk model-years/hour on %d cores' % (d['value']/1e3, d['cpu_baseline']['cores']))"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > $OUT/sanitizer_${tool}_$TAG.log 2>&1
  echo "== compute-sanitizer $tool: $(grep -c 'ERROR SUMMARY' $OUT/sanitizer_${tool}_$TAG.log) summary line(s): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_$TAG.log | tail -1); smoke: $(grep -c 'smoke OK' $OUT/sanitizer_${tool}_$TAG.log)"
done
