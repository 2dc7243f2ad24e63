#!/bin/bash
# Round 2: ncu launch list + full captures at 512 members per handle
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
M=${MEMBERS:-512}
# launch list of one model year behind 2 untimed years + warm-up year: launches per year 2496 + ... ; skip = 3 years
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
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1]))[:28]:
    print("  %-60s n=%4d avg %8.1f us  share %5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
PY
# full capture: tracer pair, BIOGEM step kernels, coupling, EMBM, barotropic solve (perturbed ensemble, 2-year-old state)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_colt|k_co_col|k_bg_step|k_tc_apply|k_tc_partial|k_embm|k_baro_blk|k_velc1" -s 4000 -c 16 \
  -o $OUT/prof_$TAG -f python tools/prof_run.py --members $M --spin 192 --steps 6 --variant col --perturb > $OUT/prof_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/details_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/source_$TAG.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/raw_$TAG.csv")))
h = rows[0]
def col(n):
    return h.index(n) if n in h else -1
names = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [col(n) for n in names]
for r in rows[2:]:
    print(" | ".join((r[i][:46] if i >= 0 and i < len(r) else "-") for i in idx))
PY
rm -f $OUT/prof_$TAG.ncu-rep
ls -la $OUT | tail -8
