"""Materialise a cGENIE job directory (namelists + input data) for cg_create().

The reference's `new-job` tool (tools/new-job.py, tools/config_utils.py:221-281)
writes data_genie / data_GOLD / data_EMBM / data_goldSIC into the job directory
and copies data/<module>/ inputs under input/<module>/.  This does the same from
the packed inputs in configs/inputs.npz (tools/pack_inputs.py), so tests and
benchmarks run where /root/reference does not exist.  Reals are written with
repr() so the parsed doubles are bit-identical to the reference files'.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = os.path.join(ROOT, "configs", "inputs.npz")

# BASELINE.json configs -> (world, levels, tracers, ocean steps / year)
CONFIGS = {
    "eb_go_gs_36x36x8": dict(world="worbe2", maxk=8, maxl=2, nyear=100),     # config #1 (t100, the CPU test job)
    "eb_go_gs_36x36x16": dict(world="worjh2", maxk=16, maxl=2, nyear=96),     # physics of configs #2-4
    "eb_go_gs_36x36x16_L16": dict(world="worjh2", maxk=16, maxl=16, nyear=96),  # + 14 passive tracers on ts
}


def _fmt(v):
    if isinstance(v, bool):
        return ".TRUE." if v else ".FALSE."
    if isinstance(v, str):
        return '"%s"' % v
    if isinstance(v, float):
        return repr(v)
    return str(v)


def _write_nml(path, group, items):
    with open(path, "w") as f:
        f.write("&%s\n" % group)
        for k, v in items:
            f.write("%s=%s,\n" % (k, _fmt(v)))
        f.write("&END\n")


def timestepping(nyear):
    """tools/config_utils.py:103-162: primary step and relative loop counts."""
    return dict(genie_timestep=3600.0 * 24.0 * 365.25 / 5.0 / nyear, katm_loop=1, ksic_loop=5, kocn_loop=5)


def materialise(jobdir, config="eb_go_gs_36x36x8", overrides=None):
    """Write the job directory; `overrides` maps namelist keys ('go_diff(1)', 'ea_rmax', 'ma_...') to values."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    ov = dict(overrides or {})
    world, maxk, maxl, nyear = cfg["world"], cfg["maxk"], cfg["maxl"], cfg["nyear"]
    z = np.load(INPUTS)
    os.makedirs(os.path.join(jobdir, "input", "goldstein"), exist_ok=True)
    os.makedirs(os.path.join(jobdir, "input", "embm"), exist_ok=True)
    os.makedirs(os.path.join(jobdir, "input", "goldsteinseaice"), exist_ok=True)
    k1 = z[world + "/k1"]
    for mod in ("goldstein", "embm", "goldsteinseaice"):
        with open(os.path.join(jobdir, "input", mod, world + ".k1"), "w") as f:
            for row in k1:
                f.write(" ".join("%2d" % x for x in row) + "\n")
    with open(os.path.join(jobdir, "input", "goldstein", world + ".psiles"), "w") as f:
        for row in z[world + "/psiles"]:
            f.write(" ".join("%3d" % int(x) for x in row) + "\n")
    with open(os.path.join(jobdir, "input", "goldstein", world + ".paths"), "w") as f:
        npi = z[world + "/npi"]
        f.write(" ".join(str(int(n)) for n in npi) + "\n")
        p = 0
        for n in npi:
            f.write("\n")
            for _ in range(int(n)):
                f.write("%d %d %d\n" % tuple(int(x) for x in z[world + "/paths"][p]))
                p += 1
    for nm, fn in (("taux_u", "taux_u.interp"), ("tauy_u", "tauy_u.interp"), ("taux_v", "taux_v.interp"),
                   ("tauy_v", "tauy_v.interp"), ("uncep", "uncep.silo"), ("vncep", "vncep.silo")):
        with open(os.path.join(jobdir, "input", "embm", fn), "w") as f:
            f.write("\n".join(repr(float(x)) for x in z["winds/" + nm]) + "\n")

    def sect(prefix, items):
        out = []
        for k, v in items:
            key = prefix + "_" + k
            out.append((k, ov.pop(key) if key in ov else v))
        for key in [k for k in ov if k.startswith(prefix + "_")]:
            out.append((key[len(prefix) + 1:], ov.pop(key)))
        return out

    ts = timestepping(nyear)
    _write_nml(os.path.join(jobdir, "data_genie"), "GENIE_CONTROL_NML", sect("ma", [
        ("flag_ebatmos", True), ("flag_goldsteinocean", True), ("flag_goldsteinseaice", True), ("flag_ents", False),
        ("flag_biogem", False), ("flag_atchem", False), ("flag_sedgem", False), ("flag_rokgem", False),
        ("flag_gemlite", False), ("katm_loop", ts["katm_loop"]), ("ksic_loop", ts["ksic_loop"]),
        ("kocn_loop", ts["kocn_loop"]), ("conv_kocn_katchem", 2), ("conv_kocn_kbiogem", 2),
        ("genie_timestep", ts["genie_timestep"]), ("genie_solar_constant", 1368.0), ("fname_topo", world),
        ("dim_GOLDSTEINNLONS", 36), ("dim_GOLDSTEINNLATS", 36), ("dim_GOLDSTEINNLEVS", maxk),
        ("dim_GOLDSTEINNTRACS", maxl)]))
    _write_nml(os.path.join(jobdir, "data_GOLD"), "INI_GOLD_NML", sect("go", [
        ("indir_name", "input/goldstein"), ("igrid", 0), ("world", world), ("ans", "n"), ("yearlen", 365.25),
        ("nyear", nyear), ("temp0", 5.0), ("temp1", 5.0), ("rel", 0.9), ("scf", 2.0), ("diff(1)", 2000.0),
        ("diff(2)", 1.0e-5), ("adrag", 2.5), ("hosing", 0.0), ("hosing_trend", 0.0), ("nyears_hosing", 0),
        ("fwanomin", "n"), ("albocn", 0.05), ("iconv", 0), ("imld", 0), ("iediff", 0), ("ieos", 0), ("dosc", True),
        ("diso", True), ("ssmaxsurf", 10.0), ("ssmaxdeep", 10.0), ("saln0", 34.9)]))
    _write_nml(os.path.join(jobdir, "data_EMBM"), "INI_EMBM_NML", sect("ea", [
        ("indir_name", "input/embm"), ("igrid", 0), ("world", world), ("xu_wstress", "taux_u.interp"),
        ("yu_wstress", "tauy_u.interp"), ("xv_wstress", "taux_v.interp"), ("yv_wstress", "tauy_v.interp"),
        ("u_wspeed", "uncep.silo"), ("v_wspeed", "vncep.silo"), ("ans", "n"), ("yearlen", 365.25), ("nyear", nyear),
        ("ndta", 5), ("scf", 2.0), ("rmax", 0.85), ("diffamp(1)", 5.0e6), ("diffamp(2)", 1.0e6), ("diffwid", 1.0),
        ("difflin", 0.1), ("betaz(1)", 0.0), ("betaz(2)", 0.4), ("betam(1)", 0.0), ("betam(2)", 0.4), ("t_co2", 0),
        ("radfor_scl_co2", 1.0), ("radfor_pc_co2_rise", 0.0), ("radfor_scl_ch4", 1.0), ("radfor_pc_ch4_rise", 0.0),
        ("radfor_scl_n2o", 1.0), ("radfor_pc_n2o_rise", 0.0), ("tatm", 10.0), ("relh0_ocean", 0.0),
        ("relh0_land", 0.0), ("extra1a", -0.03), ("extra1b", 0.17), ("extra1c", 0.18), ("scl_fwf", 1.0),
        ("z1_embm", 10.0), ("atchem_radfor", "n"), ("diffa_scl", 1.0), ("diffa_len", 0), ("dosc", True),
        ("delf2x", 5.77), ("olr_adj0", 0.0), ("olr_adj", 0.0), ("t_eqm", 12.371), ("useforc", False),
        ("orbit_radfor", "n"), ("albedop_offs", 0.20), ("albedop_amp", 0.36), ("albedop_skew", 0.0),
        ("albedop_skewp", 0), ("albedop_mod2", 0.0), ("albedop_mod4", 0.0), ("albedop_mod6", 0.0), ("orogswitch", 0),
        ("t_orog", 0), ("t_lice", 0), ("t_d18o", 0), ("par_wind_polar_avg", 0), ("par_sich_max", 9999.9),
        ("par_albsic_min", 0.2), ("par_albsic_max", 0.7)]))
    _write_nml(os.path.join(jobdir, "data_goldSIC"), "INI_SIC_NML", sect("gs", [
        ("indir_name", "input/goldsteinseaice"), ("igrid", 0), ("world", world), ("ans", "n"), ("yearlen", 365.25),
        ("nyear", nyear), ("diffsic", 2000.0), ("dosc", True), ("impsic", False), ("par_sica_thresh", 1.0),
        ("par_sich_thresh", 1000.0)]))
    if ov:
        raise KeyError("unknown namelist overrides: %s" % sorted(ov))
    return jobdir
