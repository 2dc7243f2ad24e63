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


def ref_signatures():
    """Dummy-argument lists of the reference's hot-path module procedures (name, type, rank, INTENT per argument), parsed from the
    reference's Fortran with numpy.f2py.crackfortran: tests/golden/ref_signatures.json.  tests/test_fortran_shim.py holds the shim
    modules (fortran/*_b200.f90) to them -- the drop-in must take exactly what genie_loop_wrappers.f90 passes."""
    import contextlib
    import io
    from numpy.f2py import crackfortran
    crackfortran.verbose = 0
    want = {"goldstein/goldstein.f90": ["step_goldstein"], "embm/embm.f90": ["step_embm", "surflux"],
            "goldsteinseaice/gold_seaice.f90": ["step_seaice"],
            "biogem/biogem.f90": ["biogem_forcing", "step_biogem", "biogem_tracercoupling", "biogem_climate", "biogem_climate_sol"],
            "atchem/atchem.f90": ["step_atchem", "cpl_flux_ocnatm", "cpl_comp_atmocn", "cpl_comp_embm"],
            "sedgem/sedgem.f90": ["cpl_flux_ocnsed", "cpl_comp_ocnsed"], "rokgem/rokgem.f90": ["reinit_flux_rokocn"]}

    def walk(b, out):
        if b.get("block") in ("subroutine", "function"):
            out[b["name"].lower()] = b
        for c in b.get("body", []):
            walk(c, out)

    sigs = {}
    for f, names in want.items():
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            blocks = crackfortran.crackfortran(["/root/reference/src/" + f])
        procs = {}
        for b in blocks:
            walk(b, procs)
        for n in names:
            p = procs[n]
            sigs[n] = {"file": "src/" + f, "args": [{"name": a.lower(), "type": p["vars"][a].get("typespec"),
                                                     "rank": len(p["vars"][a].get("dimension", [])),
                                                     "intent": sorted(p["vars"][a].get("intent") or [])} for a in p["args"]]}
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_signatures.json")
    json.dump({"source": "numpy.f2py.crackfortran over /root/reference/src (tools/make_golden.py ref_signatures)", "procedures": sigs},
              open(out, "w"), indent=1)


if os.path.isdir("/root/reference/src"):
    ref_signatures()
