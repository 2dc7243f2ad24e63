#!/bin/bash
# per-kernel launch durations (ncu, cold cache, serialised) of a few steady-state ocean cycles
TAG=${1:-l}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 120 --csv --log-file $OUT/launches_$TAG.csv \
  python tools/prof_run.py --members 128 --spin 400 --steps 8 --variant col > $OUT/launches_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/launches_$TAG.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][:40]
    d.setdefault(k, []).append(float(r[vi].replace(",", "")) / 1000.0)
for k, v in d.items(): print("%-42s n=%3d avg %8.1f us  total %8.1f" % (k, len(v), sum(v) / len(v), sum(v)))
PY
