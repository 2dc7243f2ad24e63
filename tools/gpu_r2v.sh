#!/bin/bash
# Round 2: ieos = 1 parity; launch list + full capture of the BIOGEM block in its packets / cells form at 512 members
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
M=${MEMBERS:-512}
timeout 900 python -m pytest tests/test_gpu_eos.py tests/test_gpu_ediff.py tests/test_gpu_parity.py -m gpu -q -s > $OUT/pytest_gpu_$TAG.log 2>&1
grep -n "passed\|failed\|FAILED\|Error\|ieos=" $OUT/pytest_gpu_$TAG.log | head -20
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 8200 -c 1600 --csv --log-file $OUT/launches_${TAG}_M$M.csv \
  python bench.py --members $M --steps 1 --warmup 1 --spinup-years 2 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches_${TAG}_M$M.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
t = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > iv:
        try: t[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")) / 1e3)
        except ValueError: pass
tot = sum(sum(v) for v in t.values())
print("launch list: %d launches, %.1f ms total" % (sum(len(v) for v in t.values()), tot / 1e3))
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1]))[:16]:
    print("  %-60s n=%4d avg %8.1f us  share %5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_bg_step|k_bg_cell|k_tc_partial" -s 300 -c 6 \
  -o $OUT/prof_$TAG -f python tools/prof_run.py --members $M --spin 192 --steps 6 --variant col --perturb > $OUT/prof_full_$TAG.log 2>&1
tail -2 $OUT/prof_full_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/details_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/source_$TAG.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/raw_$TAG.csv")))
h = rows[0]
names = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [h.index(n) if n in h else -1 for n in names]
for r in rows[2:]:
    print(" | ".join((r[i][:40] if i >= 0 and i < len(r) else "-") for i in idx))
PY
rm -f $OUT/prof_$TAG.ncu-rep
