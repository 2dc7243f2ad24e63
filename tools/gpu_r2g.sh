#!/bin/bash
# Round 2: convection variants and async release in the bench state + launch list of the bench state
TAG=${1:-r2g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_col_proof.py -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1
CG_CO_V=3 timeout 600 python -m pytest tests/test_gpu_col_proof.py tests/test_gpu_col.py -m gpu -q -s -k "not century" >> $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
bash tools/gpu_ab_col.sh $TAG "CG_X=0" "CG_CO_V=3" "CG_COL_CFG=2" "CG_COL_CFG=2 CG_CO_V=3"
# launch list of the bench's timed region (100-year-old perturbed ensemble): first launches = spin-up (249 600) + warm-up year
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 252200 -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/launches_$TAG.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][:44]
    d.setdefault(k, []).append(float(r[vi].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])): print("%-46s n=%4d avg %8.1f us  total %9.1f  %5.1f %%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
PY
