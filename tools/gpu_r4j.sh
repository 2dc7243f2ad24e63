#!/bin/bash
# Round 2, session 3: convection decisions in lockstep form (co_decide_static; CG_CO_MINB=4xx) against the default, then -- if it wins --
# the per-cell proof / tile / smoke / drift tests with it switched on
OUT=gpurun_out
mkdir -p $OUT
for envs in "CG_CO_MINB=232" "CG_CO_MINB=416" "CG_CO_MINB=412" "CG_CO_MINB=420"; do
  echo "== $envs"
  env $envs timeout 200 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4j.log
python - <<'PY' > $OUT/ab_r4j_pick.txt
import re
t = open("gpurun_out/ab_r4j.log").read()
v = dict(re.findall(r"== CG_CO_MINB=(\d+)\n.*?tstepo_flux=([0-9.]+)", t))
v = {k: float(x) for k, x in v.items()}
best = min((k for k in v if k != "232"), key=lambda k: v[k])
print(best if v[best] < v["232"] - 10.0 else "none")
PY
PICK=$(cat $OUT/ab_r4j_pick.txt)
echo "pick: $PICK" | tee -a $OUT/ab_r4j.log
if [ "$PICK" != "none" ]; then
  export CG_CO_MINB=$PICK
  timeout 400 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_tiles.py -q -x -s 2>&1 | tail -25 | tee $OUT/pytest_gpu_lockstep_r4j.log
  timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a $OUT/pytest_gpu_lockstep_r4j.log
  timeout 300 python -m pytest tests/test_gpu_col.py -q -x 2>&1 | tail -8 | tee -a $OUT/pytest_gpu_lockstep_r4j.log
fi
