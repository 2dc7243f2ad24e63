"""The oracle's chemistry against values PRINTED IN THE LITERATURE it is taken from (mitigation 3).

The reference holds no golden vectors for this path and cannot be built here, so the oracle (oracle/cgo_biogem.c, a line-by-line
restatement of src/common/gem_carbchem.f90) is anchored where an anchor exists outside both code bases: the "check values" the
DOE (1994) Handbook of methods for the analysis of the various parameters of the carbon dioxide system in sea water (Dickson &
Goyet, eds., ch. 5) and Zeebe & Wolf-Gladrow (2001, appendix A) print for S = 35, t = 25 degC, P = 0 next to each formula.
gem_carbchem.f90 works on the seawater (SWS) pH scale in mol (kg-soln)-1; the handbook's values are on the total scale, so each
comparison undoes the scale conversions the Fortran applies (gem_carbchem.f90:117-133), using the published total sulphate /
fluoride and K_S / K_F -- themselves check values.  A wrong coefficient, sign or scale conversion in the restatement moves these
numbers by far more than the tolerances."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle_lib import Oracle, lib

NAMES = "K1 K2 K KB KW KSI KHF KHSO4 KP1 KP2 KP3 KH2S KNH4 KCAL KARG QCO2 QO2".split()
T25, S35 = 298.15, 35.0


def carbconst(D, T, S):
    L = lib()
    L.cgo_test_carbconst.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)]
    L.cgo_test_carbconst.restype = None
    cc = (C.c_double * 17)()
    L.cgo_test_carbconst(D, T, S, cc)
    return dict(zip(NAMES, list(cc)))


@pytest.fixture(scope="module")
def cc():
    return carbconst(0.0, T25, S35)


@pytest.fixture(scope="module")
def scales(cc):
    """ln(total -> SWS) and ln(free -> SWS) factors at S = 35, t = 25 from the published K_S, K_F and total sulphate / fluoride."""
    ST, FT = 0.02824, 0.00007                       # mol (kg-soln)-1 at S = 35 (DOE 1994 ch. 5 table 2)
    KS = math.exp(-2.30)                            # Dickson (1990): ln K_S = -2.30, free scale
    KF_free = math.exp(-5.80) / (1.0 + ST / KS)     # Dickson & Riley (1979): ln K_F = -5.80 on the total scale
    free2tot = math.log(1.0 + ST / KS)
    free2sws = math.log(1.0 + ST / KS + FT / KF_free)
    return {"t2s": free2sws - free2tot, "f2s": free2sws, "f2t": free2tot}


def test_weiss_1974_co2_solubility(cc):
    # Weiss (1974), mol kg-1 atm-1; DOE (1994) ch. 5 eq. 3 check value: ln K0 = -3.5617
    assert abs(math.log(cc["QCO2"]) - (-3.5617)) < 1e-4


def test_mucci_1983_solubility_products(cc):
    # Mucci (1983): pK*sp(calcite) = 6.3693, pK*sp(aragonite) = 6.1883 (Zeebe & Wolf-Gladrow 2001, A.10)
    assert abs(-math.log10(cc["KCAL"]) - 6.3693) < 1e-4
    assert abs(-math.log10(cc["KARG"]) - 6.1883) < 1e-4


def test_mehrbach_refit_dickson_millero_1987(cc):
    # Dickson & Millero (1987) refit of Mehrbach et al. (1973), SWS scale: pK1 = 5.8372, pK2 = 8.9554 (their table 4 equations
    # evaluated at S = 35, 25 degC; quoted e.g. in Zeebe & Wolf-Gladrow 2001, A.1.2) -- gem_carbchem's default set
    assert abs(-math.log10(cc["K1"]) - 5.8372) < 1e-4
    assert abs(-math.log10(cc["K2"]) - 8.9554) < 1e-4


def test_doe_1994_check_values_total_scale(cc, scales):
    t2s, f2s = scales["t2s"], scales["f2s"]
    # Dickson (1990) boric acid: ln KB = -19.7964.  gem_carbchem uses the mol (kg-H2O)-1 coefficients + ln(1 - 0.001005 S),
    # which differs from the handbook's mol (kg-soln)-1 refit by 0.005 in ln KB (0.5 % in KB)
    assert abs(math.log(cc["KB"]) - t2s - (-19.7964)) < 1e-2
    # Millero (1995) water: ln KW = -30.434 (total); the SWS formula gem_carbchem carries has the constant 148.9802 where the
    # handbook's total-scale form has 148.96502
    assert abs(math.log(cc["KW"]) - (148.9802 - 148.96502) - (-30.434)) < 1e-3
    # Dickson (1990) bisulphate: ln KS = -2.30 (free scale)
    assert abs(math.log(cc["KHSO4"]) - f2s - (-2.30)) < 5e-3
    # Dickson & Riley (1979) hydrogen fluoride: ln KF = -5.80 (total scale)
    assert abs(math.log(cc["KHF"]) - t2s - (-5.80)) < 5e-3
    # Millero (1995) / Yao & Millero (1995) phosphoric acid: ln K1P = -3.71, ln K2P = -13.727, ln K3P = -20.24 (total scale;
    # the SWS fits differ by the same 0.015 in the constant term)
    assert abs(math.log(cc["KP1"]) - 0.015 - (-3.71)) < 5e-3
    assert abs(math.log(cc["KP2"]) - 0.015 - (-13.727)) < 5e-3
    assert abs(math.log(cc["KP3"]) - 0.015 - (-20.24)) < 5e-3
    # Millero (1995) silicic acid: ln KSi = -21.61
    assert abs(math.log(cc["KSI"]) - 0.015 - (-21.61)) < 5e-3


def test_oxygen_saturation(cc):
    # air-saturated O2 at S = 35, 25 degC: 206 umol kg-1 (Weiss 1970 / Garcia & Gordon 1992 tables agree to 0.5 %)
    assert abs(cc["QO2"] * 0.20946 * 1e6 - 206.0) < 1.5


def test_pressure_correction_millero_1995(cc):
    """Millero (1979, 1995) molal volume / compressibility corrections: no effect at the surface; in cold deep water (2 degC,
    4000 m) the calcite solubility product is 2.0 - 2.6 times its surface value -- the textbook reason for the lysocline (e.g.
    Zeebe & Wolf-Gladrow 2001, section 1.1.6 / fig. 1.1.7) -- and K1, K2 rise by 30 - 60 %."""
    assert carbconst(0.0, T25, S35) == cc
    top, deep = carbconst(0.0, 275.15, S35), carbconst(4000.0, 275.15, S35)
    assert 2.0 < deep["KCAL"] / top["KCAL"] < 2.6
    assert 1.3 < deep["K1"] / top["K1"] < 1.6 and 1.2 < deep["K2"] / top["K2"] < 1.6


def test_carbonate_solve_is_an_equilibrium(cc):
    """sub_calc_carb at surface-ocean values: the solution satisfies the mass balance and the two mass-action laws it was not
    built from directly, and lands where every carbonate-system calculator does for these inputs (pH_SWS 8.0 - 8.1, fCO2
    330 - 400 uatm for DIC 2000 / TA 2300 umol kg-1 at 25 degC, S = 35)."""
    L = lib()
    L.cgo_test_calc_carb.argtypes = [C.c_double] * 6 + [C.POINTER(C.c_double)] * 2
    L.cgo_test_calc_carb.restype = C.c_int
    ccv = (C.c_double * 17)(*[cc[n] for n in NAMES])
    carb = (C.c_double * 10)()
    carb[0] = 10.0 ** -7.8
    DIC, ALK = 2000e-6, 2300e-6
    assert L.cgo_test_calc_carb(DIC, ALK, 1.028e-2, 0.0, 0.0, S35, ccv, carb) == 0
    H, co2, co3, hco3, fug = carb[0], carb[1], carb[2], carb[3], carb[4]
    assert abs((co2 + hco3 + co3) / DIC - 1.0) < 1e-12
    assert abs(H * hco3 / co2 / cc["K1"] - 1.0) < 1e-6 and abs(H * co3 / hco3 / cc["K2"] - 1.0) < 1e-6
    assert 8.0 < -math.log10(H) < 8.1 and 330e-6 < fug < 400e-6
    assert 4.5 < carb[5] < 5.5                     # calcite saturation state of warm surface water


def test_isotope_notation_round_trip():
    L = lib()
    for f in (L.cgo_test_iso_delta, L.cgo_test_iso_fraction):
        f.restype = C.c_double
    L.cgo_test_iso_delta.argtypes = [C.c_double] * 3
    L.cgo_test_iso_fraction.argtypes = [C.c_double] * 2
    std13 = 0.011202     # gem_cmn.f90:630; the accepted VPDB 13C/12C ratio is 0.0112372 (Craig 1957) -- cGENIE's own value
    fr = L.cgo_test_iso_fraction(-6.5, std13)
    assert abs(fr / (1.0 - fr) / std13 - (1.0 - 6.5e-3)) < 1e-15
    assert abs(L.cgo_test_iso_delta(1.0, fr, std13) - (-6.5)) < 1e-10


def test_insolation_annual_global_mean():
    """radfor (embm.f90:2383-2522): the area- and time-mean top-of-atmosphere insolation of any orbit is S0 / 4 / sqrt(1 - e^2)
    (Berger 1978); with S0 = 1368 W m-2 and today's e = 0.0167 that is 342.05 W m-2.  The sine-latitude grid has equal areas."""
    o = Oracle("worbe2", maxk=8, maxl=2, nyear=100)
    sol = o.f("solfor")
    assert sol.size == 36 * 100
    assert abs(sol.mean() - 1368.0 / 4.0 / math.sqrt(1.0 - 0.0167 ** 2)) < 0.7
    lat_mean = sol.reshape(100, 36).mean(axis=0)       # solfor(maxj, nyear), Fortran order
    assert lat_mean[0] < 200.0 and lat_mean[-1] < 200.0 and 390.0 < lat_mean[17] < 430.0      # ~175 W m-2 at the poles, ~417 at the equator
    assert np.all(sol >= 0.0)


def test_oracle_constants_are_the_references():
    """Every physical constant the oracle carries as a macro (oracle/cgo_impl.h, cgo_biogem.c) against the value of the reference's own
    PARAMETER declaration, evaluated from the reference's text (goldstein_lib.f90, embm_lib.f90, gem_cmn.f90: tests/golden/
    ref_constants.json, tools/make_golden.py) -- bit for bit, derived constants (rhosc, rfluxsc, conv_yr_s ...) included."""
    import json
    import math
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "ref_constants.json")))["constants"]
    assert len(ref) >= 59
    defs = {}
    for f in ("cgo_impl.h", "cgo_biogem.c"):
        for line in open(os.path.join(root, "oracle", f)):
            m = re.match(r"#define\s+((?:CG|BG)_[A-Z0-9_]+)\s+(.+?)\s*(?:/\*.*)?$", line)
            if m:
                defs.setdefault(m.group(1), m.group(2))

    def value(name, depth=0):
        assert depth < 20
        expr = re.sub(r"\b((?:CG|BG)_[A-Z0-9_]+)\b", lambda q: repr(value(q.group(1), depth + 1)), defs[name])
        return float(eval(expr, {"__builtins__": {}}, {"atan": math.atan}))

    for macro, r in ref.items():
        assert macro in defs, macro
        got = value(macro)
        assert got.hex() == r["hex"], "%s = %r, the reference's %s (%s) = %r" % (macro, got, r["name"], r["file"], r["value"])


def test_gas_exchange_tables_are_the_references():
    """Schmidt-number and Bunsen-coefficient rows of the four gases of the frozen selection, as the oracle (cgo_biogem.c) and the product
    (csrc/cg_biogem.cu) carry them, against src/common/gem_data.f90:69-136 (tests/golden/ref_gas_tables.json)."""
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "ref_gas_tables.json")))["tables"]
    label = {"IA_PCO2": "pCO2", "IA_PO2": "pO2", "IA_PCFC11": "pCFC11", "IA_PCFC12": "pCFC12"}
    for path in (os.path.join(root, "oracle", "cgo_biogem.c"), os.path.join(root, "cgenie_b200", "csrc", "cg_biogem.cu")):
        text = open(path).read()
        for arr, key, n in (("Sc", "schmidt", 4), ("Bu", "bunsen", 6)):
            body = text[text.index("double %s[][%d] = {" % (arr, n + 1)):]
            body = body[:body.index("};")]
            rows = re.findall(r"\{(IA_\w+),([^}]*)\}", body)
            assert len(rows) == 4, (path, arr)
            for ia, nums in rows:
                got = [float(x) for x in nums.split(",")]
                assert got == ref[key][label[ia]], (path, arr, ia, got, ref[key][label[ia]])


def test_restoring_forcing_is_the_references():
    """The atmospheric restoring forcing of the frozen configuration -- which tracers, time constant, two-point signal, end members 0 / 1
    -- as the job directory (cgenie_b200/jobdir.py BG_RESTORE) and the oracle (cgo_biogem.c, sub_init_force_restore_atm block) carry it,
    against data/biogem/worjh2_preindustrial of the reference (tests/golden/ref_forcing_worjh2_preindustrial.json)."""
    import json
    import os
    import re
    from cgenie_b200 import jobdir
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, "tests", "golden", "ref_forcing_worjh2_preindustrial.json")))
    assert ref["n_atm_rows"] == 19
    assert sorted(jobdir.BG_RESTORE) == ref["restored"]
    for ia, (tc, val) in jobdir.BG_RESTORE.items():
        r = ref["restore_atm"][str(ia)]
        assert jobdir.BG_ATM_NAMES[ia] == r["name"]
        assert tc == r["tconst"] and not r["flux"]
        assert r["sig"] == [[0.0, val], [999999.0, val]]
        assert r["I"] == [0.0, 0.0, 1296] and r["II"] == [0.0, 1.0, 1296]
    text = open(os.path.join(root, "oracle", "cgo_biogem.c")).read()
    body = text[text.index("static const double sig[][2] = {"):]
    body = body[:body.index("};")]
    rows = dict(re.findall(r"\{(IA_\w+),\s*([^}]+)\}", body))
    assert len(rows) == len(ref["restored"])
    for r in ref["restore_atm"].values():
        assert float(rows["IA_" + r["name"].upper()]) == r["sig"][0][1]
    assert "b->rst_tconst[la] = 0.1;" in text
