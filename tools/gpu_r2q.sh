#!/bin/bash
# Round 2: iediff parity + `ncu --set full` of the top kernels at 512 members (perturbed ensemble, 2-year-old state)
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
M=${MEMBERS:-512}
timeout 900 python -m pytest tests/test_gpu_ediff.py tests/test_gpu_parity.py tests/test_gpu_hosing.py -m gpu -q -s --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error\|iediff=" $OUT/pytest_gpu_$TAG.log | head -20
# matching launches per ocean step: flux, co, embm, baro, velc1 (5) + per BIOGEM block: surface, sweep, apply, 2 partial (5)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_colt|k_co_col|k_bg_step|k_tc_apply|k_tc_partial|k_embm|k_baro_blk|k_velc1" -s 1400 -c 15 \
  -o $OUT/prof_$TAG -f python tools/prof_run.py --members $M --spin 192 --steps 6 --variant col --perturb > $OUT/prof_full_$TAG.log 2>&1
tail -3 $OUT/prof_full_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/details_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/source_$TAG.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/raw_$TAG.csv")))
h = rows[0]
names = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [h.index(n) if n in h else -1 for n in names]
print(" | ".join(n[-28:] for n in names)); print(" | ".join((rows[1][i] if i >= 0 else "-") for i in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:40] if i >= 0 and i < len(r) else "-") for i in idx))
PY
rm -f $OUT/prof_$TAG.ncu-rep
