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


def ref_date_files():
    """The two netCDF files the reference holds (data/main/main_restart_0.nc, main_fluxes_0_date.nc: 220 bytes each, written by the
    netCDF library for genie-main) as hex fixtures: tests/golden/ref_*.hex.  They pin the netCDF-3 codec (header layout, padding,
    begin offsets, big-endian INT data) to the real library's output: tests/test_restart_nc.py reproduces them byte for byte."""
    import binascii
    for n in ("main_restart_0.nc", "main_fluxes_0_date.nc"):
        with open("/root/reference/data/main/" + n, "rb") as fh:
            d = fh.read()
        with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_" + n + ".hex"), "w") as fh:
            fh.write(binascii.hexlify(d).decode() + "\n")


if os.path.isdir("/root/reference/data/main"):
    ref_date_files()
