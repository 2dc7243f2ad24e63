#!/bin/bash
# Round 2: packets / cells form of the BIOGEM sweep
TAG=${1:-r2u}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_z_packets_cells.py tests/test_gpu_biogem.py tests/test_gpu_z_slice.py tests/test_gpu_z_sig.py tests/test_gpu_zz_series_year.py tests/test_gpu_restart.py -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1
tail -15 $OUT/pytest_gpu_$TAG.log
bash tools/gpu_ab.sh $TAG "CG_BG_PD=1" "CG_BG_PD=0" "CG_BG_PD=1 CG_BG_CELL_MINB=4" "CG_BG_PD=1 CG_BG_PK_MINB=4" "CG_BG_PD=1 CG_BG_PK_MINB=2"
