#!/bin/bash
# Round 2, session 3: imld = 1 under BIOGEM on the device + the BIOGEM tests around the edited kernels, then a short bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_mld.py tests/test_gpu_biogem.py tests/test_gpu_z_packets_cells.py tests/test_gpu_z_sig2.py tests/test_gpu_z_slice.py -q -x -s 2>&1 | tail -40 | tee $OUT/pytest_gpu_mld_bg_r4g.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_r4g.json 2> $OUT/bench_r4g.err
python -c "
import json
d=json.load(open('$OUT/bench_r4g.json')); r=d['roofline']
print('bench: value %.3fM ms/yr %.2f e2e %.3fM tstepo %.1f us frac %.3f traffic %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, 1e3*r['avg_launch_ms'], r['frac'], r['traffic']))
for k,v in r['other_families'].items(): print('   %-9s %.3f ms/call' % (k, v['avg_call_ms']))
" | tee -a $OUT/pytest_gpu_mld_bg_r4g.log
