#!/bin/bash
# Round 2, session 3: convection kernel with the averaging form fixed at compile time (regions: 126 / 96 / 80 / 72 / 64 registers, no spills)
OUT=gpurun_out
mkdir -p $OUT
for envs in "CG_CO_MINB=216" "CG_CO_MINB=220" "CG_CO_MINB=224" "CG_CO_MINB=228" "CG_CO_MINB=232" "CG_CO_MINB=116"; do
  echo "== $envs"
  env $envs timeout 240 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4d.log
