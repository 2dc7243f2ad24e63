#!/bin/bash
# A/B of the coupling sums taken one block ahead (CG_TC_AHEAD, default on; cg_run and the per-module path) on one B200:
# schedule bit-identity tests first, then the default bench line without the CPU arm, alternating the knob.
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_col.py -q -x -k "concurrent_schedule or module_by_module" > $OUT/pytest_tc_$1.log 2>&1
tail -3 $OUT/pytest_tc_$1.log
for v in 1 0 1 0; do
  CG_TC_AHEAD=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('CG_TC_AHEAD=$v value %.4f M ms %.3f e2e %.4f M launches %d blown %d' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, j['gpu_launches'], j['blown_up_members']))" | tee -a $OUT/ab_tc_$1.log
done
