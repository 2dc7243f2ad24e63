#!/bin/bash
# Round 2, session 3: convection kernel, blocks of the poleward rows first (CG_CO_POLAR=1) against row order
OUT=gpurun_out
mkdir -p $OUT
for envs in "CG_CO_POLAR=0" "CG_CO_POLAR=1"; do
  echo "== $envs"
  env $envs timeout 240 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4e.log
