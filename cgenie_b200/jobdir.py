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
    # multi-island topographies of the reference's data/goldstein (2 and 3 islands): the barotropic closure with matmult
    "eb_go_gs_p0055c_36x36x16": dict(world="p0055c", maxk=16, maxl=2, nyear=96),
    "eb_go_gs_p0251a_36x36x16": dict(world="p0251a", maxk=16, maxl=2, nyear=96),
    # configs #2-4: BIOGEM + ATCHEM with the frozen 16-tracer selection (DESIGN.md "BIOGEM configuration")
    "eb_go_gs_ac_bg_36x36x16": dict(world="worjh2", maxk=16, maxl=16, nyear=96, biogem=True),
}

# frozen BIOGEM configuration: selected tracers (ids of data/main/tracer_define.*) and their initial values
BG_OCN = {1: 0.0, 2: 0.0, 3: 2.244E-03, 4: 0.4, 5: -150.0, 8: 2.159E-06, 10: 1.696E-04, 12: 2.363E-03, 15: 0.0, 16: 0.0,
          17: 0.0, 20: 0.0, 35: 1.025E-02, 45: 0.0, 46: 0.0, 50: 5.282E-02}
BG_SED = (3, 4, 5, 8, 14, 15, 16, 33, 34)
BG_ATM = {1: 0.0, 2: 0.0, 3: 278.0E-06, 4: -6.5, 5: 0.0, 6: 0.2095, 18: 0.0, 19: 0.0}
BG_ATM_NAMES = {3: "pCO2", 4: "pCO2_13C", 5: "pCO2_14C", 6: "pO2", 18: "pCFC11", 19: "pCFC12"}
# restoring forcing of data/biogem/worjh2_preindustrial: tracer -> (time constant / yr, constant signal value)
BG_RESTORE = {3: (0.1, 2.780000E-04), 4: (0.1, -6.50), 5: (0.1, 38.4), 18: (0.1, 0.0), 19: (0.1, 0.0)}


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


def _materialise_biogem(jobdir, z, sect):
    """data_GEM / data_BIOGEM / data_ATCHEM, the prescribed wind speed and the atmospheric restoring forcing
    (formats: gem_util.f90:356-419, 511-536; biogem_data.f90:1346-1357; biogem_lib.f90:1421-1481)."""
    bdir = os.path.join(jobdir, "input", "biogem")
    fdir = os.path.join(bdir, "forcing")
    os.makedirs(fdir, exist_ok=True)
    with open(os.path.join(bdir, "windspeed.dat"), "w") as f:
        for row in z["biogem/worjh2_windspeed"]:
            f.write(" ".join(repr(float(x)) for x in row) + "\n")
    with open(os.path.join(fdir, "configure_forcings_atm.dat"), "w") as f:
        f.write("-START-OF-DATA-\n")
        for ia in range(1, 20):
            tc, _ = BG_RESTORE.get(ia, (1.0, 0.0))
            f.write(" %02d  %s  %s  F  F  F\n" % (ia, "T" if ia in BG_RESTORE else "F", repr(tc)))
        f.write("-END-OF-DATA-\n")
    wet = z["worjh2/k1"][1:37, 1:37] <= 16   # file rows j = maxj..1
    for ia, (_, val) in BG_RESTORE.items():
        base = os.path.join(fdir, "biogem_force_restore_atm_" + BG_ATM_NAMES[ia])
        for tag, fill in (("_I", 0.0), ("_II", 1.0)):
            with open(base + tag + ".dat", "w") as f:
                for row in wet:
                    f.write(" ".join(repr(fill if w else 0.0) for w in row) + "\n")
        with open(base + "_sig.dat", "w") as f:
            f.write("-START-OF-DATA-\n0.0 %s\n999999.0 %s\n-END-OF-DATA-\n" % (repr(val), repr(val)))
    _write_nml(os.path.join(jobdir, "data_GEM"), "INI_GEM_NML", sect("gm",
        [("ocn_select(%d)" % i, i in BG_OCN) for i in range(1, 96)] + [("sed_select(%d)" % i, i in BG_SED) for i in range(1, 80)] +
        [("atm_select(%d)" % i, i in BG_ATM) for i in range(1, 20)] +
        [("par_carbconstset_name", "Mehrbach"), ("par_carbchem_pH_tolerance", 0.001), ("par_carbchem_pH_iterationmax", 100),
         ("ctrl_carbchem_fail", True)]))
    _write_nml(os.path.join(jobdir, "data_BIOGEM"), "INI_BIOGEM_NML", sect("bg",
        [("ocn_init(%d)" % i, float(v)) for i, v in sorted(BG_OCN.items())] +
        [("par_misc_t_start", 0.0), ("par_misc_t_runtime", 1001.0), ("ctrl_misc_t_BP", False), ("ctrl_misc_Snorm", True),
         ("par_misc_brinerejection_frac", 0.0), ("ctrl_force_sed_closedsystem", True), ("ctrl_force_GOLDSTEInTS", True),
         ("ctrl_force_GOLDSTEInTSonly", False), ("ctrl_force_seaice", False), ("ctrl_force_windspeed", True),
         ("par_gastransfer_a", 0.310), ("par_indir_name", "input/biogem"), ("par_fordir_name", "input/biogem/forcing"),
         ("par_windspeed_file", "windspeed.dat"), ("ctrl_force_oldformat", True), ("par_bio_prodopt", "1N1T_PO4MM"),
         ("par_bio_k0_PO4", 2.0E-06), ("par_bio_c0_PO4", 0.050E-06), ("par_bio_red_POP_PON", 16.0),
         ("par_bio_red_POP_POC", 106.0), ("par_bio_red_POP_PO2", -138.0), ("par_bio_red_PON_ALK", -1.00),
         ("par_bio_red_DOMfrac", 0.66), ("par_bio_red_RDOMfrac", 0.0), ("opt_bio_CaCO3toPOCrainratio", "Ridgwelletal2007ab"),
         ("par_bio_red_POC_CaCO3", 0.2), ("par_bio_red_POC_CaCO3_pP", 0.0), ("par_bio_remin_DOMlifetime", 0.5),
         ("ctrl_bio_remin_POC_fixed", True), ("par_bio_remin_fun", "efolding"), ("ctrl_bio_remin_POC_ballast", False),
         ("par_bio_remin_POC_frac2", 0.05), ("par_bio_remin_POC_eL1", 500.0), ("par_bio_remin_POC_eL2", 1000000.0),
         ("par_bio_remin_POC_dfrac2", 0.0), ("par_bio_remin_POC_c0frac2", 0.1E-6), ("ctrl_bio_remin_CaCO3_fixed", True),
         ("par_bio_remin_CaCO3_frac2", 0.5), ("par_bio_remin_CaCO3_eL1", 1000.0), ("par_bio_remin_CaCO3_eL2", 1000000.0),
         ("par_bio_remin_sinkingrate", 125.0), ("par_bio_remin_k_O2", 1.0), ("par_bio_remin_c0_O2", 8.0E-6),
         ("par_d13C_DIC_Corg_ef", 25.0), ("par_Fgeothermal", 0.0), ("ctrl_bio_preformed", False)]))
    _write_nml(os.path.join(jobdir, "data_ATCHEM"), "INI_ATCHEM_NML", sect("ac",
        [("atm_init(%d)" % i, float(v)) for i, v in sorted(BG_ATM.items())] + [("par_atm_F14C", 0.0)]))


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

    biogem = bool(cfg.get("biogem", False))
    if biogem:
        _materialise_biogem(jobdir, z, sect)
    ts = timestepping(nyear)
    _write_nml(os.path.join(jobdir, "data_genie"), "GENIE_CONTROL_NML", sect("ma", [
        ("flag_ebatmos", True), ("flag_goldsteinocean", True), ("flag_goldsteinseaice", True), ("flag_ents", False),
        ("flag_biogem", biogem), ("flag_atchem", biogem), ("flag_sedgem", False), ("flag_rokgem", False),
        ("flag_gemlite", False), ("katm_loop", ts["katm_loop"]), ("ksic_loop", ts["ksic_loop"]),
        ("kocn_loop", ts["kocn_loop"]), ("conv_kocn_katchem", 2), ("conv_kocn_kbiogem", 2),
        ("genie_timestep", ts["genie_timestep"]), ("genie_solar_constant", 1368.0), ("fname_topo", world),
        ("dim_GOLDSTEINNLONS", 36), ("dim_GOLDSTEINNLATS", 36), ("dim_GOLDSTEINNLEVS", maxk),
        ("dim_GOLDSTEINNTRACS", maxl)]))
    _write_nml(os.path.join(jobdir, "data_GOLD"), "INI_GOLD_NML", sect("go", [
        ("indir_name", "input/goldstein"), ("igrid", 0), ("world", world), ("ans", "n"), ("yearlen", 365.25),
        ("nyear", nyear), ("temp0", 5.0), ("temp1", 5.0), ("rel", 0.9), ("scf", 2.0), ("diff(1)", 2000.0),
        ("diff(2)", 1.0e-5), ("adrag", 2.5), ("hosing", 0.0), ("hosing_trend", 0.0), ("nyears_hosing", 0),
        ("fwanomin", "n"), ("albocn", 0.05), ("iconv", 0), ("imld", 0), ("mldpebuoycoeff", 0.15), ("mldketaucoeff", 2.5),
        ("mldwindkedec", 25.0), ("iediff", 0), ("ieos", 0), ("dosc", True),
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
