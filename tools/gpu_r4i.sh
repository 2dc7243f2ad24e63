#!/bin/bash
# Round 2, session 3: convection kernel with two member tiles per block (64 threads): 56 / 48 / 40 registers = 36 / 40 / 42 warps per SM
OUT=gpurun_out
mkdir -p $OUT
for envs in "CG_CO_MINB=318" "CG_CO_MINB=320" "CG_CO_MINB=321"; do
  echo "== $envs"
  env $envs timeout 240 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4i.log
