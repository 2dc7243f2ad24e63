#!/bin/bash
# Round 2, session 3: A/B of the register-capped convection kernel (CG_CO_MINB) and the tiled tracer-window flux kernel (CG_COL_WIN)
# on the bench state (512 members, perturbed, 100-year-old ocean); CUDA-event us per ocean step per family + a hash of the final ts
OUT=gpurun_out
mkdir -p $OUT
for envs in "CG_X=0" "CG_CO_MINB=16" "CG_CO_MINB=20" "CG_CO_MINB=24" "CG_COL_WIN=1" "CG_COL_WIN=2" "CG_CO_MINB=20 CG_COL_WIN=1"; do
  echo "== $envs"
  env $envs timeout 240 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4a.log
