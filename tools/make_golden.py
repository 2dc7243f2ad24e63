#!/usr/bin/env python
"""Self-generated regression vector for the oracle (NOT a reference-generated golden: the reference
cannot be built in this image).  One model year of BASELINE config #1."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle  # noqa: E402
from test_oracle import inventory  # noqa: E402

o = Oracle("worbe2", maxk=8, maxl=2, nyear=100)
o.run(500)
inv = inventory(o, 8, 2)
vals = dict(T=float(inv[0]), S=float(inv[1]), tq_sum=float(o.f("tq").sum()), psi_min=float(o.f("psi").min()),
            psi_max=float(o.f("psi").max()), ice=float(o.f("varice").sum()), cost=float(o.f("cost").sum()))
out = os.path.join(ROOT, "tests", "golden", "oracle_eb_go_gs_36x36x8_1yr.json")
json.dump({"source": "oracle/ (self-generated, parity unpinned)", "config": "worbe2 36x36x8 L=2 nyear=100, 500 koverall",
           "values": vals}, open(out, "w"), indent=1)
print(vals)
