#!/bin/bash
# Round 2, last session (r4): whole GPU suite, smoke, default bench line + reference arm, launch list and full captures at the bench
# state (512 members, 100-year-old ocean), compute-sanitizer over smoke()
TAG=${1:-r4f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -12 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open('$OUT/bench_$TAG.json')); r=d['roofline']
print('bench: value %.3fM ms/yr %.2f e2e %.3fM (serial %.3fM, h2d %.0f MB) tstepo %.1f us frac %.3f cpu %.1fk' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['serial_value']/1e6, d['e2e']['h2d_bytes_per_step']/1e6, 1e3*r['avg_launch_ms'], r['frac'], d['cpu_baseline']['value']/1e3))
for k,v in r['other_families'].items(): print('   %-9s %.3f ms/call  %.0f GB/s  frac %.3f' % (k, v['avg_call_ms'], v['achieved'], v['frac']))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
python -c "
import json
d=json.load(open('$OUT/bench_ref_$TAG.json')); print('reference arm: %.1fk model-years/hour on %d cores' % (d['value']/1e3, d['cpu_baseline']['cores']))"
for C in 1 2 5; do
  timeout 600 python bench.py --config $C --steps 5 --warmup 3 > $OUT/bench_config${C}_$TAG.json 2> $OUT/bench_config${C}_$TAG.err
  python -c "
import json
d=json.load(open('$OUT/bench_config${C}_$TAG.json')); r=d['roofline']; print('config $C: value %.4g %s, ms/step %.3f, tstepo %.1f us, %.1f GB/s (frac %.3f)' % (d['value'], d['unit'], d['ms_per_step'], 1e3*r['avg_launch_ms'], r['achieved'], r['frac']))" || tail -3 $OUT/bench_config${C}_$TAG.err
done
# launch list of the timed year in the bench state (100 untimed years + warm-up year skipped: 2497 launches per model year)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 252900 -c 2600 --csv --log-file $OUT/launches_${TAG}_M512.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-serial > $OUT/bench_under_ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
try:
    rows = list(csv.reader(open("$OUT/launches_${TAG}_M512.csv")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    t = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > iv:
            try: t[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")) / 1e3)
            except ValueError: pass
    tot = sum(sum(v) for v in t.values())
    print("launch list (bench state): %d launches, %.1f ms total" % (sum(len(v) for v in t.values()), tot / 1e3))
    for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1]))[:22]:
        print("  %-60s n=%4d avg %8.1f us  share %5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
except Exception as ex:
    print("launch list failed:", ex)
PY
# full capture at the bench state: per 2 ocean steps 2 x (flux, convection) + surface, packets, cells = 7 matching launches
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_colt|k_co_col|k_bg_step|k_bg_cell|k_embm|k_baro_blk4" -s 52800 -c 22 \
  -o $OUT/prof_$TAG -f python tools/prof_run.py --members 512 --spin 9600 --steps 6 --variant col --perturb > $OUT/prof_full_$TAG.log 2>&1
tail -2 $OUT/prof_full_$TAG.log
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/details_$TAG.csv 2>/dev/null
python - <<PY
import csv
try:
    rows = list(csv.reader(open("$OUT/raw_$TAG.csv")))
    h = rows[0]
    names = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [h.index(n) if n in h else -1 for n in names]
    for r in rows[2:]:
        print(" | ".join((r[i][:40] if i >= 0 and i < len(r) else "-") for i in idx))
except Exception as ex:
    print("raw page failed:", ex)
PY
rm -f $OUT/prof_$TAG.ncu-rep
# smoke() under memcheck (racecheck / initcheck of smoke: clean in r3f, its kernels are unchanged), the kernels new in this session under all three
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > $OUT/sanitizer_memcheck_$TAG.log 2>&1
echo "== compute-sanitizer memcheck: $(grep 'ERROR SUMMARY' $OUT/sanitizer_memcheck_$TAG.log | tail -1); smoke: $(grep -c 'smoke OK' $OUT/sanitizer_memcheck_$TAG.log)"
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_new.py > $OUT/sanitizer_new_${tool}_$TAG.log 2>&1
  echo "== compute-sanitizer $tool (mixed layer, extended series): $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitizer_new_${tool}_$TAG.log | tail -1); ok: $(grep -c 'sanitize_new OK' $OUT/sanitizer_new_${tool}_$TAG.log)"
done
