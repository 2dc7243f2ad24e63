#!/bin/bash
# Round 2, session 3: extended time-series integrals on the device (new test + the neighbouring series / BIOGEM tests), then the A/B of
# the decisions-only convection kernel (CG_CO_V=4 | 5) against the default on the bench state
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_z_sig2.py tests/test_gpu_z_sig.py tests/test_gpu_zz_series_year.py tests/test_gpu_biogem.py tests/test_gpu_z_packets_cells.py -q -x -s 2>&1 | tail -60 | tee $OUT/pytest_gpu_sig2_r4c.log
for envs in "CG_X=0" "CG_CO_V=4" "CG_CO_V=5"; do
  echo "== $envs"
  env $envs timeout 240 python tools/prof_run.py --members 512 --spin 9600 --steps 96 --variant col --perturb --profile --hash 2>&1 | tail -3
done 2>&1 | tee $OUT/ab_r4c.log
