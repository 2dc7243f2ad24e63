#!/bin/bash
# lazy per-module cycle (CG_NOLAZY=1: off): bit-identity tests, then the e2e composition and the bench line with / without
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_col.py tests/test_gpu_parity.py -q -x -k "module_by_module or concurrent_schedule or momentum_embm_seaice or full_model" > $OUT/pytest_lazy_$1.log 2>&1
tail -3 $OUT/pytest_lazy_$1.log
timeout 200 python tools/e2e_split.py 2>&1 | tail -8 | tee $OUT/e2e_split_$1.log
CG_NOLAZY=1 timeout 200 python tools/e2e_split.py 2>&1 | tail -8 | sed 's/^/NOLAZY /' | tee -a $OUT/e2e_split_$1.log
for v in 0 1; do
  CG_NOLAZY=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('CG_NOLAZY=$v value %.4f M ms %.3f e2e %.4f M launches %d blown %d' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, j['gpu_launches'], j['blown_up_members']))" | tee -a $OUT/ab_lazy_$1.log
done
