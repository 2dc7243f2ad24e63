#!/usr/bin/env python
"""Append the tracer-step DRAM traffic of an `ncu --set full` raw page (ncu -i x.ncu-rep --page raw --csv) to profiles/ncu_traffic.json,
the table bench.py's roofline.traffic is read from.  usage: ncu_traffic_add.py raw.csv ROUND MEMBERS "state text" [source path]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, rnd, members, state = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
src = sys.argv[5] if len(sys.argv) > 5 else raw
rows = list(csv.reader(open(raw)))
h, u = rows[0], rows[1]
ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tscale = {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
acc = {}
for r in rows[2:]:
    k = r[ik].split("(")[0].replace("void ", "").replace("cg::", "")
    if not (k.startswith("k_tstep_col") or k.startswith("k_co_")):
        continue
    acc.setdefault(k, []).append((float(r[ir].replace(",", "")) * scale[u[ir]], float(r[iw].replace(",", "")) * scale[u[iw]],
                                  float(r[it].replace(",", "")) * tscale[u[it]]))
kern, total = {}, 0.0
for k, v in acc.items():
    n = len(v)
    rd, wr, t = (sum(x[q] for x in v) / n for q in range(3))
    kern[k] = {"dram_read": int(rd), "dram_write": int(wr), "time_us_under_ncu": round(t, 1), "launches_averaged": n}
    total += rd + wr
p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(p))
d["captures"].append({"round": rnd, "config": "eb_go_gs_ac_bg_36x36x16", "members": members, "variant": "col", "state": state, "kernels": kern,
                      "dram_bytes_per_tstepo_launch": int(total), "algorithmic_bytes_per_tstepo_launch": 12511 * 288 * members,
                      "source": src})
json.dump(d, open(p, "w"), indent=1)
print("tstepo traffic %.3f GB per launch pair = %.3f x algorithmic" % (total / 1e9, total / (12511 * 288 * members)))
