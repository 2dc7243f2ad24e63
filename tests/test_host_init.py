"""CPU tests: the C-ABI loads, every declared symbol exists, and the product's host-side
initialisation (cg_create: namelist + data-file parsing, grid, masks, drag, barotropic
factorisation, island solves, insolation table ...) is BIT-IDENTICAL to the oracle's restatement of
initialise_goldstein / initialise_embm / initialise_seaice.  No compute call, no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cgenie_b200 import _lib, materialise
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [("eb_go_gs_36x36x8", dict(world="worbe2", maxk=8, maxl=2, nyear=100)),
         ("eb_go_gs_36x36x16_L16", dict(world="worjh2", maxk=16, maxl=16, nyear=96)),
         # topographies with two and three islands (unit island solves, erisl after matinv_gold)
         ("eb_go_gs_p0055c_36x36x16", dict(world="p0055c", maxk=16, maxl=2, nyear=96)),
         ("eb_go_gs_p0251a_36x36x16", dict(world="p0251a", maxk=16, maxl=2, nyear=96))]


def test_header_symbols_exported(built):
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "cgenie_b200.h")).read()
    names = set(re.findall(r"\b(cg_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 35
    for n in sorted(names):
        assert hasattr(lib, n), "library does not export " + n
    assert set(_lib.SYMBOLS) == names


def test_create_fails_loudly_on_missing_job(built, tmp_path):
    lib = _lib.load()
    h = _lib.P()
    rc = lib.cg_create(str(tmp_path).encode(), 1, 0, C.byref(h))
    assert rc == 2 and b"could not open" in lib.cg_last_error()


def test_out_of_scope_option_rejected(built, tmp_path):
    lib = _lib.load()
    # an unknown mixed-layer option, a second equation-of-state option and a spatially varying ediff grid are not on the device path
    for q, ov in enumerate(({"go_imld": 2}, {"go_ieos": 2}, {"go_iediff": 1, "go_ediffvar": 0.5}, {"go_iediff": 3})):
        d = tmp_path / ("job%d" % q)
        materialise(str(d), "eb_go_gs_36x36x8", ov)
        h = _lib.P()
        rc = lib.cg_create(str(d).encode(), 1, 0, C.byref(h))
        assert rc == 3 and b"outside the B200 hot path" in lib.cg_last_error() or b"on the B200 hot path" in lib.cg_last_error(), ov
        assert rc == 3, ov
    # ... the ones that are: accepted
    for q, ov in enumerate(({"go_ieos": 1}, {"go_iconv": 1}, {"go_iediff": 2, "go_ediff0": 0.275e-4}, {"go_imld": 1})):
        d = tmp_path / ("ok%d" % q)
        materialise(str(d), "eb_go_gs_36x36x8", ov)
        h = _lib.P()
        assert lib.cg_create(str(d).encode(), 1, 0, C.byref(h)) == 0, (ov, lib.cg_last_error())
        lib.cg_destroy(h)


class HostOnly:
    """cg_create without cg_initialise: constants only."""

    def __init__(self, jobdir):
        self.L = _lib.load()
        self.h = _lib.P()
        rc = self.L.cg_create(jobdir.encode(), 1, 0, C.byref(self.h))
        assert rc == 0, self.L.cg_last_error()

    def const(self, name):
        n = self.L.cg_const_size(self.h, name.encode())
        assert n >= 0, name
        out = np.empty(n)
        assert self.L.cg_get_const(self.h, name.encode(), 0, out.ctypes.data_as(_lib.D), n) == 0
        return out

    def iconst(self, name):
        n = self.L.cg_const_size(self.h, name.encode())
        assert n >= 0, name
        out = np.empty(n, dtype=np.int32)
        assert self.L.cg_get_iconst(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int32)), n) == 0
        return out

    def close(self):
        self.L.cg_destroy(self.h)


def same(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.flatnonzero(a.view(np.uint64 if a.dtype == np.float64 else a.dtype) !=
                         b.view(np.uint64 if b.dtype == np.float64 else b.dtype))
    assert bad.size == 0, "%s differs at %d places, first %d: %r vs %r" % (what, bad.size, bad[0], a[bad[0]], b[bad[0]])


@pytest.mark.parametrize("config,okw", CASES)
def test_host_init_bit_exact_vs_oracle(built, tmp_path, config, okw):
    materialise(str(tmp_path), config)
    p = HostOnly(str(tmp_path))
    o = Oracle(**okw)
    I = J = 36
    K, L = okw["maxk"], okw["maxl"]
    # 1-D metrics: product arrays are indexed with the Fortran index like the oracle's
    for name, n in [("ds", J + 1), ("dsv", J), ("rds2", J), ("s", J + 1), ("c", J + 1), ("sv", J + 1), ("cv", J + 1),
                    ("rc", J + 1), ("rc2", J + 1), ("rcv", J), ("rdsv", J), ("cv2", J), ("rds", J + 1), ("asurf", J + 1),
                    ("dz", K + 1), ("dza", K + 1), ("zro", K + 1), ("zw", K + 1), ("rdz", K + 1), ("rdza", K + 1)]:
        same(p.const(name)[:n], o.f(name)[:n], name)
    same(p.const("ssmax")[1:K], o.f("ssmax")[1:K], "ssmax")
    for name in ("k1", "ku", "mk", "iroff", "jroff"):
        same(p.iconst(name), o.i(name), name)
    same(p.iconst("getj") != 0, o.i("getj") != 0, "getj")
    for name in ("ips", "ipf", "ias", "iaf"):
        same(p.iconst(name)[1:J + 1], o.i(name)[1:J + 1], name)
    assert p.iconst("jsf")[0] == int(o.s("jsf")) and p.iconst("ntot")[0] == int(o.s("ntot"))
    for name in ("rh", "drag", "rtv", "rtv3", "rhosing", "gap", "ratm", "ubisl", "psisl", "erisl", "diffa", "albcl", "ca",
                 "pmeadj", "solfor", "us_dztau", "us_dztav"):
        same(p.const(name), o.f(name), name)
    sc = p.const("scalars")
    names = ["dphi", "rdphi", "dzz", None, "diff1", "diff2", "adrag", "ec1", "ec2", "ec3", "ec4", "rpmesco", "rsictscsf",
             "dtatm", "rdtdim", "rfluxsca", "rpmesca", "dtsic", "sic_rdtdim", "diffsic"]
    for v, n in zip(sc, names):
        if n:
            same(np.array([v]), np.array([o.s(n)]), n)
    same(np.array([sc[3]]), o.f("dt")[1:2], "dt")
    # initial state
    ts_o = o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :]
    same(p.const("ts0").reshape(K, J, I, L)[..., :2], ts_o[..., :2], "ts0")
    same(p.const("rho0").reshape(K, J, I), o.f("rho").reshape(K + 1, J + 2, I + 2)[1:, 1:J + 1, 1:I + 1], "rho0")
    same(p.const("tq0"), o.f("tq"), "tq0")
    # winds: the oracle holds uatm after initialise_embm; the product keeps the same array
    same(p.const("uatm"), o.f("uatm"), "uatm")
    p.close()
    o.close()


def test_mixed_layer_tables_bit_exact_vs_oracle(built, tmp_path):
    """imld = 1: the wind energy input (goldstein.f90:112-141, from tau = scf * stress) and its decay with depth (:1675-1686) as the
    product's host initialisation builds them against the oracle's, for a non-default scf and decay scale."""
    kw = dict(imld=1, scf=1.7, mldketaucoeff=3.0, mldwindkedec=40.0)
    materialise(str(tmp_path), "eb_go_gs_36x36x8", {"go_" + k: v for k, v in kw.items()})
    p = HostOnly(str(tmp_path))
    o = Oracle("worbe2", maxk=8, maxl=2, nyear=100, **kw)
    o.run(5)                                   # mldketau is formed in step_goldstein
    same(p.const("mldketau"), o.f("mldketau"), "mldketau")
    same(p.const("mlddec")[1:9], o.f("mlddec")[1:9], "mlddec")
    same(p.const("mlddecd")[1:9], o.f("mlddecd")[1:9], "mlddecd")
    assert p.const("mldketau").max() > 0
    p.close()
    o.close()


def test_jobdir_roundtrip_bit_exact(tmp_path):
    """ASCII written by materialise() parses back to the packed doubles (repr round trip)."""
    materialise(str(tmp_path), "eb_go_gs_36x36x8")
    z = np.load(os.path.join(ROOT, "configs", "inputs.npz"))
    back = np.array([float(t) for t in open(tmp_path / "input" / "embm" / "taux_u.interp").read().split()])
    assert np.array_equal(back.view(np.uint64), z["winds/taux_u"].view(np.uint64))


def test_product_constants_are_the_references():
    """The product's own constants (csrc/cg_host.hpp, cg_biogem.hpp: written separately from the oracle's macros) against the values of
    the reference's PARAMETER declarations (tests/golden/ref_constants.json, evaluated from the reference's text), bit for bit."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_constants.json")))["constants"]
    defs = {}
    for f in ("cg_host.hpp", "cg_biogem.hpp"):
        text = open(os.path.join(ROOT, "cgenie_b200", "csrc", f)).read()
        for stmt in re.findall(r"constexpr double ([^;]+);", text):
            for part in re.split(r",\s*(?=k[A-Z])", stmt):
                name, expr = part.split("=", 1)
                defs[name.strip()] = expr.strip()

    def value(name, depth=0):
        assert depth < 20
        expr = re.sub(r"\bk[A-Z][A-Za-z0-9]*\b", lambda q: repr(value(q.group(0), depth + 1)), defs[name])
        return float(eval(expr, {"__builtins__": {}}, {}))

    names = {"CG_USC": "kUsc", "CG_RSC": "kRsc", "CG_DSC": "kDsc", "CG_FSC": "kFsc", "CG_GSC": "kGsc", "CG_RH0SC": "kRh0sc",
             "CG_RHOSC": "kRhosc", "CG_TSC": "kTsc", "CG_CPSC": "kCpsc", "CG_RHOAIR": "kRhoair", "CG_RHO0": "kRho0", "CG_RHOAO": "kRhoao",
             "CG_M2MM": "kM2mm", "CG_MM2M": "kMm2m", "CG_RFLUXSC": "kRfluxsc", "CG_CPA": "kCpa", "CG_CONST1": "kConst1",
             "CG_CONST2": "kConst2", "CG_CONST3": "kConst3", "CG_CONST4": "kConst4", "CG_CONST5": "kConst5", "CG_SIGMA": "kSigma",
             "CG_EMO": "kEmo", "CG_EMA": "kEma", "CG_TFREEZ": "kTfreez", "CG_HLV": "kHlv", "CG_HLF": "kHlf", "CG_HLS": "kHls",
             "CG_CONSIC": "kConsic", "CG_ZEROC": "kZeroc", "CG_CPO_ICE": "kCpoIce", "CG_RHOICE": "kRhoice", "CG_HMIN": "kHmin",
             "CG_RHMIN": "kRhmin", "CG_RHOOI": "kRhooi", "CG_RHOIO": "kRhoio", "CG_RRHOLF": "kRrholf", "CG_CO20": "kCo20",
             "CG_CH40": "kCh40", "CG_N2O0": "kN2o0", "CG_ALPHACH4": "kAlphaCh4", "CG_ALPHAN2O": "kAlphaN2o", "CG_TSIC": "kTsic",
             "CG_CD": "kCd", "BG_ZEROC": "kBgZeroC", "BG_NULL": "kBgNull", "BG_NULLSMALL": "kBgNullSmall", "BG_YR_S": "kBgYrS",
             "BG_LAMBDA_14C": "kBgLambda14C", "BG_M3_KG": "kBgM3Kg", "BG_PI": "kBgPi", "BG_REARTH": "kBgREarth"}
    for macro, k in names.items():
        assert k in defs, k
        assert value(k).hex() == ref[macro]["hex"], (k, value(k), ref[macro])


@pytest.mark.parametrize("config", ["eb_go_gs_36x36x8", "eb_go_gs_ac_bg_36x36x16"])
def test_job_namelists_are_the_reference_defaults_but_for_the_configuration(tmp_path, config):
    """The base configs of the BASELINE jobs live in the un-vendored cgenie-data repository, so the job directories of this repo are
    reconstructions.  What can be held to the reference is held: every key a job namelist carries exists in the reference's
    own *-defaults.nml (tests/golden/ref_namelist_defaults.json, parsed from src/), and every value IS the reference's default except
    the ones that make the configuration -- module flags, time stepping, topography / dimensions, tracer selections, initial
    inventories, the forcing directory."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_namelist_defaults.json")))["defaults"]
    materialise(str(tmp_path), config)
    allowed = {"data_genie": r"flag_\w+|k\w+_loop|conv_kocn_k\w+|genie_timestep|fname_topo|dim_goldsteinn\w+",
               "data_GOLD": r"world|nyear", "data_EMBM": r"world|nyear", "data_goldSIC": r"world|nyear",
               "data_GEM": r"(ocn|sed|atm)_select\(\d+\)", "data_BIOGEM": r"ocn_init\(\d+\)|par_fordir_name", "data_ATCHEM": r"atm_init\(\d+\)"}

    def parse(path):
        out = {}
        for ln in open(path):
            ln = ln.strip()
            if not ln or ln[0] in "&/!" or "=" not in ln:
                continue
            k, v = ln.rstrip(",").split("=", 1)
            out[k.strip().lower()] = v.strip()
        return out

    def norm(v):
        v = v.strip().strip('"').strip("'")
        if v.upper() in (".TRUE.", "T", ".T."):
            return True
        if v.upper() in (".FALSE.", "F", ".F."):
            return False
        try:
            return float(v.lower().replace("d", "e"))
        except ValueError:
            return v

    nkeys = ndev = 0
    for f, pat in allowed.items():
        p = tmp_path / f
        if not p.exists():
            assert f in ("data_GEM", "data_BIOGEM", "data_ATCHEM") and "ac_bg" not in config
            continue
        for k, v in parse(p).items():
            nkeys += 1
            assert k in ref[f], "%s: key %s is not in the reference's defaults file" % (f, k)
            if norm(v) != norm(ref[f][k]):
                ndev += 1
                assert re.fullmatch(pat, k), "%s: %s = %s differs from the reference's default %s" % (f, k, v, ref[f][k])
    print("%s: %d keys, %d set by the configuration" % (config, nkeys, ndev))
    assert nkeys > 100 and ndev < 70


def test_builtin_parameter_defaults_are_the_references():
    """The defaults the product (struct Params, csrc/cg_host.hpp) and the oracle (PD / PI_ in oracle/cgo_driver.c) fall back to when a
    namelist does not carry a key, against the reference's *-defaults.nml (tests/golden/ref_namelist_defaults.json)."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_namelist_defaults.json")))["defaults"]
    pool = {}
    for f in ("data_GOLD", "data_EMBM", "data_goldSIC", "data_genie", "data_BIOGEM"):
        for k, v in ref[f].items():
            pool.setdefault(k, []).append(v)
    alias = {"diff1": "diff(1)", "diff2": "diff(2)", "diffamp1": "diffamp(1)", "diffamp2": "diffamp(2)", "betaz1": "betaz(1)",
             "betaz2": "betaz(2)", "betam1": "betam(1)", "betam2": "betam(2)", "solconst": "genie_solar_constant",
             "z1_embm": "z1_embm", "maxi": "dim_goldsteinnlons", "maxj": "dim_goldsteinnlats", "maxk": "dim_goldsteinnlevs",
             "maxl": "dim_goldsteinntracs"}

    def num(v):
        v = v.strip().strip('"').strip("'")
        if v.upper() in (".TRUE.", "T"):
            return 1.0
        if v.upper() in (".FALSE.", "F"):
            return 0.0
        if v.lower() in ("y", "n"):               # the reference's CHARACTER switches (atchem_radfor, fwanomin ...)
            return 1.0 if v.lower() == "y" else 0.0
        try:
            return float(v.lower().replace("d", "e"))
        except ValueError:
            return v

    mine = {}
    hpp = open(os.path.join(ROOT, "cgenie_b200", "csrc", "cg_host.hpp")).read()
    body = hpp[hpp.index("struct Params {"):hpp.index("bool set(const std::string &name, double v);")]
    body = re.sub(r"//[^\n]*", "", body)
    for name, val in re.findall(r"\b([a-z_][a-z0-9_]*)\s*=\s*(-?[0-9][0-9.eE+-]*|true|false)\b", body):
        mine.setdefault(name, []).append(("product", 1.0 if val == "true" else 0.0 if val == "false" else float(val)))
    drv = open(os.path.join(ROOT, "oracle", "cgo_driver.c")).read()
    for name, val in re.findall(r"\bP(?:D|I_)\(\s*([a-z_0-9]+)\s*,\s*(-?[0-9][0-9.eE+-]*)\s*\)", drv):
        mine.setdefault(name, []).append(("oracle", float(val)))
    # what makes a configuration, not a default: set by every job directory (timestepping() of jobdir.py, config_utils.py:103-162)
    config_keys = {"kocn_loop", "ksic_loop", "katm_loop", "conv_kocn_kbiogem", "conv_kocn_katchem", "genie_timestep", "maxk", "maxl",
                   "flag_biogem", "flag_atchem"}
    checked = 0
    for name, lst in mine.items():
        key = alias.get(name, name)
        if key not in pool or name in config_keys:
            continue
        want = {num(v) for v in pool[key]}
        for who, val in lst:
            assert val in want, "%s default of %s = %r, the reference's %s" % (who, name, val, sorted(map(str, want)))
            checked += 1
    assert checked >= 120, checked


@pytest.mark.skipif(not os.path.isdir("/root/reference/data/goldstein"), reason="the reference tree is not mounted here")
def test_packed_inputs_are_the_reference_data_files():
    """configs/inputs.npz (what tests, smoke() and bench.py read where /root/reference does not exist) against the reference's data
    files themselves -- topographies, island paths, wind stresses and speeds, BIOGEM's wind-speed field -- value for value.  Runs in
    the build container only (the GPU box has no reference tree)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pack_inputs", os.path.join(ROOT, "tools", "pack_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.collect()
    z = np.load(os.path.join(ROOT, "configs", "inputs.npz"))
    assert set(z.files) == set(fresh), sorted(set(z.files) ^ set(fresh))
    for k, v in fresh.items():
        assert z[k].dtype == v.dtype and z[k].shape == v.shape, k
        assert np.array_equal(z[k].view(np.uint8), np.ascontiguousarray(v).view(np.uint8)), k
