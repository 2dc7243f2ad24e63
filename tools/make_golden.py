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


# oracle macro -> (reference file under src/, PARAMETER name); the values themselves come from the reference's text
REF_CONSTANTS = {
    "goldstein/goldstein_lib.f90": dict(CG_USC="usc", CG_RSC="rsc", CG_DSC="dsc", CG_FSC="fsc", CG_GSC="gsc", CG_RH0SC="rh0sc", CG_RHOSC="rhosc",
                                        CG_TSC="tsc", CG_CPSC="cpsc", CG_RHOAIR="rhoair", CG_RHO0="rho0", CG_RHOAO="rhoao", CG_M2MM="m2mm",
                                        CG_MM2M="mm2m", CG_RFLUXSC="rfluxsc", CG_CPA="cpa", CG_PI="pi", CG_ZEROC="zeroc", CG_CPO_ICE="cpo_ice"),
    "embm/embm_lib.f90": dict(CG_CONST1="const1", CG_CONST2="const2", CG_CONST3="const3", CG_CONST4="const4", CG_CONST5="const5",
                              CG_SIGMA="sigma", CG_EMO="emo", CG_EMA="ema", CG_TFREEZ="tfreez", CG_HLV="hlv", CG_HLF="hlf", CG_HLS="hls",
                              CG_CONSIC="consic", CG_RHOICE="rhoice", CG_HMIN="hmin", CG_RHMIN="rhmin", CG_RHOOI="rhooi", CG_RHOIO="rhoio",
                              CG_RRHOLF="rrholf", CG_CO20="co20", CG_CH40="ch40", CG_N2O0="n2o0", CG_ALPHACH4="alphach4",
                              CG_ALPHAN2O="alphan2o", CG_TSIC="tsic", CG_CD="cd"),
    "common/gem_cmn.f90": dict(BG_PI="const_pi", BG_REARTH="const_rearth", BG_M3_KG="conv_m3_kg", BG_ZEROC="const_zeroc", BG_NULL="const_real_null",
                               BG_NULLSMALL="const_real_nullsmall", BG_YR_S="conv_yr_s", BG_ATM_MOL="conv_atm_mol", BG_R="const_r",
                               BG_R_SI="const_r_si", BG_V="const_v", BG_CP="const_cp", BG_CONC_MG="const_conc_mg",
                               BG_CONC_MGTOCA="const_conc_mgtoca", BG_LAMBDA_14C="const_lambda_14c", BG_PA_ATM="conv_pa_atm"),
}


def fortran_parameters(path):
    """REAL / INTEGER PARAMETER constants of a Fortran source as {lower-case name: value}: the declarations are evaluated in order, with
    the names defined so far in scope (E / D exponents, ATAN, ** as written); what cannot be evaluated (arrays, kind expressions) is
    skipped.  -fdefault-real-8 (platforms/LINUX:7-17) makes every REAL literal a double, as a Python float is."""
    import math
    import re
    text = open(path, errors="replace").read()
    text = re.sub(r"&\s*\n\s*&?", " ", text)
    ns = {"atan": math.atan, "sqrt": math.sqrt, "exp": math.exp, "log": math.log, "real": float}
    out = {}
    for line in text.split("\n"):
        line = line.split("!")[0]
        m = re.match(r"\s*(REAL|INTEGER)\b[^:]*\bPARAMETER\b[^:]*::(.*)", line, flags=re.I)
        if not m:
            continue
        depth, cur, parts = 0, "", []
        for ch in m.group(2):
            if ch == "(":
                depth += 1
            if ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        for p in parts:
            if "=" not in p:
                continue
            name, expr = p.split("=", 1)
            name, expr = name.strip().lower(), expr.strip().lower()
            if "(" in name:
                continue
            expr = re.sub(r"(\d\.?\d*|\.\d+)d([-+]?\d+)", r"\1e\2", expr)
            expr = re.sub(r"_\w+\b", "", expr) if re.search(r"\d_\w+", expr) else expr
            try:
                out[name] = float(eval(expr, {"__builtins__": {}}, dict(ns, **out)))
            except Exception:
                pass
    return out


def ref_constants():
    """tests/golden/ref_constants.json: the reference's own values of every physical constant the oracle carries as a macro."""
    vals = {}
    for f, table in REF_CONSTANTS.items():
        pars = fortran_parameters("/root/reference/src/" + f)
        for macro, name in table.items():
            if name in pars:
                vals[macro] = {"file": "src/" + f, "name": name, "value": pars[name], "hex": float(pars[name]).hex()}
            else:
                print("ref_constants: %s not found in %s" % (name, f))
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_constants.json")
    json.dump({"source": "PARAMETER declarations of /root/reference/src (tools/make_golden.py ref_constants)", "constants": vals},
              open(out, "w"), indent=1)
    return vals


if os.path.isdir("/root/reference/src"):
    ref_constants()


REF_NAMELISTS = {"data_genie": "main-defaults.nml", "data_GOLD": "goldstein/goldstein-defaults.nml", "data_EMBM": "embm/embm-defaults.nml",
                 "data_goldSIC": "goldsteinseaice/goldsteinseaice-defaults.nml", "data_GEM": "gem-defaults.nml",
                 "data_BIOGEM": "biogem/biogem-defaults.nml", "data_ATCHEM": "atchem/atchem-defaults.nml"}


def parse_namelist(path):
    out = {}
    for ln in open(path, errors="replace"):
        ln = ln.strip()
        if not ln or ln[0] in "&/!" or "=" not in ln:
            continue
        k, v = ln.rstrip(",").split("=", 1)
        out[k.strip().lower()] = v.strip()
    return out


def ref_namelist_defaults():
    """tests/golden/ref_namelist_defaults.json: the reference's *-defaults.nml files (src/), key by key, for the namelists a job directory
    of this repo carries (cgenie_b200/jobdir.py).  tests/test_host_init.py holds the reconstructed BASELINE configurations to them."""
    d = {f: parse_namelist("/root/reference/src/" + r) for f, r in REF_NAMELISTS.items()}
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_namelist_defaults.json")
    json.dump({"source": "src/*-defaults.nml of /root/reference (tools/make_golden.py ref_namelist_defaults)", "files": REF_NAMELISTS,
               "defaults": d}, open(out, "w"), indent=0)


if os.path.isdir("/root/reference/src"):
    ref_namelist_defaults()


def ref_tracer_definitions():
    """tests/golden/ref_tracer_define.json: data/main/tracer_define.{ocn,atm,sed} of the reference (name, index, dependency, type, long
    name, units per tracer) -- what the frozen 16 / 8 / 9 tracer selection of the product is held to."""
    import re
    out = {}
    for kind in ("ocn", "atm", "sed"):
        rows, on = [], False
        for ln in open("/root/reference/data/main/tracer_define." + kind, errors="replace"):
            if "-START-OF-DATA-" in ln:
                on = True
                continue
            if "-END-OF-DATA-" in ln:
                break
            if not on or not ln.strip():
                continue
            m = re.match(r"\s*(\S+)\s+(\d+)\s+(\d+)\s+(\d+)\s+'([^']*)'\s+'([^']*)'", ln)
            if m:
                rows.append({"name": m.group(1), "index": int(m.group(2)), "dep": int(m.group(3)), "type": int(m.group(4)),
                             "long_name": m.group(5), "units": m.group(6)})
        out[kind] = rows
    p = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_tracer_define.json")
    json.dump({"source": "data/main/tracer_define.{ocn,atm,sed} of /root/reference (tools/make_golden.py ref_tracer_definitions)",
               "tracers": out}, open(p, "w"), indent=0)
    return out


if os.path.isdir("/root/reference/data/main"):
    ref_tracer_definitions()


def ref_gas_tables():
    """tests/golden/ref_gas_tables.json: the Schmidt-number and Bunsen-coefficient tables of src/common/gem_data.f90:69-136, row by row
    (keyed by the row's comment label)."""
    import re
    text = open("/root/reference/src/common/gem_data.f90", errors="replace").read()
    out = {}
    for key, start in (("schmidt", "par_Sc_coef(:,:) = reshape"), ("bunsen", "par_bunsen_coef(:,:) = reshape")):
        body = text[text.index(start):]
        body = body[:body.index("/), &")]
        rows = {}
        for ln in body.split("\n"):
            m = re.match(r"\s*&\s*((?:\s*-?\d+\.\d+\s*,?)+)\s*&\s*!\s*(\S+)", ln)
            if m:
                rows[m.group(2)] = [float(x) for x in m.group(1).replace(" ", "").rstrip(",").split(",")]
        out[key] = rows
    p = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_gas_tables.json")
    json.dump({"source": "src/common/gem_data.f90:69-136 of /root/reference (tools/make_golden.py ref_gas_tables)", "tables": out},
              open(p, "w"), indent=0)
    return out


if os.path.isfile("/root/reference/src/common/gem_data.f90"):
    ref_gas_tables()


def ref_forcing_set():
    """tests/golden/ref_forcing_worjh2_preindustrial.json: the restoring forcing of data/biogem/worjh2_preindustrial as the frozen
    configuration uses it -- which atmosphere tracers are restored and with which time constant (configure_forcings_atm.dat), the
    time signal of each (*_sig.dat) and the range of the two spatial end members (*_I.dat, *_II.dat)."""
    import re
    d = "/root/reference/data/biogem/worjh2_preindustrial/"
    rows, on = [], False
    for ln in open(d + "configure_forcings_atm.dat", errors="replace"):
        if "-START-OF-DATA-" in ln:
            on = True
            continue
        if "-END-OF-DATA-" in ln:
            break
        if on and ln.strip():
            t = ln.split()
            rows.append({"restore": t[0].lower() == "t", "tconst": float(t[1]), "flux": t[2].lower() == "t"})
    names = {3: "pCO2", 4: "pCO2_13C", 5: "pCO2_14C", 6: "pO2", 18: "pCFC11", 19: "pCFC12"}
    out = {}
    for ia, r in enumerate(rows, start=1):
        if not r["restore"]:
            continue
        n = names[ia]
        sig = [[float(x) for x in ln.split()] for ln in open(d + "biogem_force_restore_atm_%s_sig.dat" % n) if re.match(r"\s*-?\d", ln)]
        ends = {}
        for e in ("I", "II"):
            v = [float(x) for x in open(d + "biogem_force_restore_atm_%s_%s.dat" % (n, e)).read().split()]
            ends[e] = [min(v), max(v), len(v)]
        out[str(ia)] = {"name": n, "tconst": r["tconst"], "flux": r["flux"], "sig": sig, "I": ends["I"], "II": ends["II"]}
    p = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_forcing_worjh2_preindustrial.json")
    json.dump({"source": "data/biogem/worjh2_preindustrial of /root/reference (tools/make_golden.py ref_forcing_set)", "restore_atm": out,
               "n_atm_rows": len(rows), "restored": sorted(int(k) for k in out)}, open(p, "w"), indent=0)
    return out


if os.path.isdir("/root/reference/data/biogem/worjh2_preindustrial"):
    ref_forcing_set()
