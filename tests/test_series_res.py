"""BIOGEM's .res time series: the Fortran edit descriptors of cg_biogem_series_write (csrc/cg_series.cpp) against hand-worked
known answers of sub_data_save_runtime's formats (biogem_data_ascii.f90:669-935), and the oracle's window integrals
(cgo_biogem_sig_update, biogem.f90:2836-2917) against numpy on the oracle's own state."""
import numpy as np

from cgenie_b200.series import write_series
from oracle_lib import Oracle

I = J = 36
K = L = 16
LA = 8


def test_res_files_formats(built, tmp_path):
    out = tmp_path / "biogem"
    write_series(str(out))
    head = open(out / "biogem_series_ocn_DIC.res").read()
    assert head == " % time (yr) / global DIC (mol) / global DIC (mol kg-1) / surface DIC (mol kg-1) / benthic DIC (mol kg-1)\n"
    assert open(out / "biogem_series_ocn_temp.res").read() == " % time (yr) / temperature (C) / _surT (C) / _benT (degrees C)\n"
    assert open(out / "biogem_series_ocn_DIC_13C.res").read().startswith(" % time (yr) / global DIC_13C (mol) / global DIC_13C (o/oo) / surface")
    assert open(out / "biogem_series_atm_pCO2.res").read() == " % time (yr) / global pCO2 (mol) / global pCO2 (atm)\n"
    assert open(out / "biogem_series_atm_humidity.res").read() == " % time (yr) / surface humidity (???)\n"
    # a window of 2 years: integrals = 2 x the means
    sig = np.zeros(3 + 3 * L + LA)
    sig[0], sig[1], sig[2] = 2.0, 2.0 * 1.35e21, 2.0 * 3.0e19
    ocn, sur, ben, atm = sig[3:3 + L], sig[3 + L:3 + 2 * L], sig[3 + 2 * L:3 + 3 * L], sig[3 + 3 * L:]
    ocn[0], sur[0], ben[0] = 2.0 * 276.65, 2.0 * 291.4, 2.0 * 272.9          # K
    ocn[1], sur[1], ben[1] = 2.0 * 34.9, 2.0 * 34.61234567, 2.0 * 34.95
    ocn[2], sur[2], ben[2] = 2.0 * 2.244e-3, 2.0 * 2.0123456e-3, 2.0 * 2.3e-3
    R = 0.011202 * (1.0 + 0.4 / 1000.0)
    ocn[3] = 2.0 * 2.244e-3 * R / (1.0 + R)                                   # d13C = +0.4
    sur[3] = 2.0 * 2.0123456e-3 * (0.011202 * 1.002) / (1.0 + 0.011202 * 1.002)  # +2.0
    ben[3] = 0.0                                                               # R = 0 -> -1000
    ocn[5] = 2.0 * 2.159e-6
    ocn[6] = -2.0 * 1.0e-101                                                   # three-digit exponent, negative
    atm[0], atm[1], atm[2] = 2.0 * 12.5, 2.0 * 0.0085, 2.0 * 278.0e-6
    Ra = 0.011202 * (1.0 - 6.5 / 1000.0)
    atm[3] = 2.0 * 278.0e-6 * Ra / (1.0 + Ra)
    write_series(str(out), sig, t_yr=0.5)
    write_series(str(out), sig, t_yr=12345.678)
    lines = open(out / "biogem_series_ocn_temp.res").read().split("\n")
    assert lines[1] == "       0.500    3.500000   18.250000   -0.250000"
    assert lines[2].startswith("   12345.678") and lines[3] == ""
    assert open(out / "biogem_series_ocn_sal.res").read().split("\n")[1] == "       0.500   34.900000   34.612346   34.950000"
    assert open(out / "biogem_series_ocn_DIC.res").read().split("\n")[1] == \
        "       0.500  0.3029400E+19  0.2244000E-02  0.2012346E-02  0.2300000E-02"
    l13 = open(out / "biogem_series_ocn_DIC_13C.res").read().split("\n")[1]
    assert l13[:12] == "       0.500" and l13[27:] == "       0.400       2.000   -1000.000" and l13[12:27].endswith("E+17")
    assert open(out / "biogem_series_ocn_PO4.res").read().split("\n")[1] == \
        "       0.500  0.2914650E+16  0.2159000E-05  0.0000000E+00  0.0000000E+00"
    assert open(out / "biogem_series_ocn_O2.res").read().split("\n")[1][12:42] == " -0.1350000E-79 -0.1000000-100"
    # an isotope whose bulk tracer is zero: const_nulliso
    assert open(out / "biogem_series_ocn_DOM_C_13C.res").read().split("\n")[1] == \
        "       0.500  0.0000000E+00    -999.999    -999.999    -999.999"
    assert open(out / "biogem_series_atm_temp.res").read().split("\n")[1] == "       0.500   12.500000"
    assert open(out / "biogem_series_atm_pCO2.res").read().split("\n")[1] == "       0.500  0.4918376E+17  0.2780000E-03"
    la = open(out / "biogem_series_atm_pCO2_13C.res").read().split("\n")[1]
    assert la[27:] == "        -6.500" and len(la) == 12 + 15 + 14
    # nothing integrated: nothing written (biogem.f90:3119)
    write_series(str(out), np.zeros_like(sig), t_yr=3.0)
    assert len(open(out / "biogem_series_ocn_temp.res").read().split("\n")) == 4


def test_oracle_sig_integrals():
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    o.biogem_setup()
    dtyr = float(2 * 5) * (3600.0 * 24.0 * 365.25 / 5.0 / 96) / (3600.0 * 24.0 * 365.25)
    want = np.zeros(3 + 3 * L + LA)
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
    for blk in range(1, 4):
        o.run(10)
        o.L.cgo_biogem_sig_update(o.h, 1000.0)
        M = np.where(wet, o.f("bg_M").reshape(K, J, I), 0.0)
        ocn = o.f("ocn").reshape(K, J, I, L)
        want[0] += dtyr
        want[1] += dtyr * M.sum()
        want[2] += dtyr * M[K - 1].sum()
        want[3:3 + L] += dtyr * (M[..., None] * ocn).reshape(-1, L).sum(axis=0) / M.sum()
    got = o.f("bg_sig")
    assert abs(got[0] - 3 * dtyr) < 1e-15
    assert np.allclose(got[1:3 + L], want[1:3 + L], rtol=1e-13, atol=0)
    t = got[0]
    # initial uniform concentrations barely move in 30 steps: the means are the initial values to a few per mil
    assert abs(got[3 + 2] / t - 2.244e-3) < 2e-5 and abs(got[3 + 5] / t - 2.159e-6) < 1e-7
    # surface (ice-free) and benthic means: bounded by the extremes of the levels they are taken from
    sur, ben, atm = got[3 + L:3 + 2 * L] / t, got[3 + 2 * L:3 + 3 * L] / t, got[3 + 3 * L:] / t
    ocn = o.f("ocn").reshape(K, J, I, L)
    top = ocn[K - 1][k1 <= K]
    assert top[:, 0].min() - 1.0 < sur[0] < top[:, 0].max() + 1.0 and sur[0] > ben[0]          # warm surface, cold abyss
    assert ocn[..., 5][wet].min() <= ben[5] <= ocn[..., 5][wet].max() * 1.01 and ben[5] > sur[5]   # PO4 depleted at the surface
    assert abs(atm[2] - 278.0e-6) < 3e-6 and abs(atm[0] - o.f("sfcatm1").reshape(J, I, LA)[..., 0].mean()) < 5.0


class _FakeEngine:
    """Records the diagnostic calls; the integrals it hands back are those of a constant ocean."""

    def __init__(self):
        self.t, self.updates, self.resets = 0.0, 0, 0

    def biogem_sig_update(self, dts, ben_Dmin):
        self.t = self.t + dts / (3600.0 * 24.0 * 365.25)
        self.updates += 1

    def biogem_sig_reset(self):
        self.t = 0.0
        self.resets += 1

    def get(self, name, member):
        assert name == "bg_sig"
        sig = np.zeros(3 + 3 * L + LA)
        sig[0], sig[1] = self.t, self.t * 1.35e21
        sig[3:3 + L] = self.t * 2.0e-3
        sig[3] = self.t * 277.15
        return sig


def test_series_windows(built, tmp_path):
    """sub_init_data_save + the window tests of diag_biogem_timeseries: a 3-year run from year 100 with yearly windows gives
    three lines centred on x.5, each integrating the 48 BIOGEM steps of its year; an explicit list of save times (as in
    biogem_save_sig.dat) with 0.5-year windows integrates 24 steps around each listed time that fits into the run."""
    from cgenie_b200.series import SeriesSaver
    nyear, kb = 96, 10
    genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / nyear
    tick = int(round(1000.0 * genie_timestep))
    dts = float(kb) * genie_timestep
    e = _FakeEngine()
    s = SeriesSaver(e, tmp_path / "a", t_runtime=3.0, t_start=100.0, sig_dt=1.0)
    assert [round(x, 9) for x in s.sig] == [0.5, 1.5, 2.5] and s.sig_i == 3
    per_window, last = [], 0
    for k in range(kb, 3 * 5 * nyear + 1, kb):
        s.step(dts, k * tick)
        if len(s.saved) > len(per_window):
            per_window.append(e.updates - last)
            last = e.updates
    assert s.saved == [100.5, 101.5, 102.5] and per_window == [48, 48, 48] and s.sig_i == 0
    lines = open(tmp_path / "a" / "biogem_series_ocn_temp.res").read().split("\n")
    assert len(lines) == 5
    assert [l[:12] for l in lines[1:4]] == ["     100.500", "     101.500", "     102.500"]
    assert lines[1][12:24] == "    4.000000"
    assert open(tmp_path / "a" / "biogem_series_ocn_DIC.res").read().split("\n")[3][12:42] == "  0.2700000E+19  0.2000000E-02"
    # explicit save times in model years; 99.0 lies before the run
    e = _FakeEngine()
    s = SeriesSaver(e, tmp_path / "b", t_runtime=3.0, t_start=100.0, sig_dt=0.5, save_times=[99.0, 100.5, 101.75, 102.9])
    counts, last = [], 0
    for k in range(kb, 3 * 5 * nyear + 1, kb):
        s.step(dts, k * tick)
        if len(s.saved) > len(counts):
            counts.append(e.updates - last)
            last = e.updates
    assert s.saved == [100.5, 101.75] and counts == [24, 24]
    # the window around 102.9 opens at 102.65 and is still filling when the run ends: without a forced save it is dropped
    assert s.sig_i == 1 and 0.0 < s.int_t_sig < 0.5 and e.updates == 48 + round(0.35 * 48)


def test_call_point_in_the_oracle():
    """genie.f90:401-405 takes the diagnostic behind biogem_climate_wrapper (:387) and ahead of cpl_flux_ocnatm_wrapper / the
    ATCHEM step (:411, :446-455): the ocean rows of that sample are those of a sample taken behind the whole block, the
    atmosphere rows (sfcatm1: cpl_comp_atmocn / cpl_comp_EMBM run behind ATCHEM) those of a sample behind the PREVIOUS block."""
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    o.biogem_setup()
    o.run(100)
    o.L.cgo_biogem_sig_update(o.h, 1000.0)           # behind block 100
    prev = o.f("bg_sig").copy()
    o.f("bg_sig")[:] = 0.0
    o.L.cgo_biogem_sig_auto(o.h, 1, 1000.0)          # inside block 110, at the reference's call point
    o.run(10)
    at = o.f("bg_sig").copy()
    o.L.cgo_biogem_sig_auto(o.h, 0, 1000.0)
    o.f("bg_sig")[:] = 0.0
    o.L.cgo_biogem_sig_update(o.h, 1000.0)           # behind block 110
    behind = o.f("bg_sig").copy()
    n_ocn = 3 + 3 * L
    assert np.array_equal(at[:n_ocn], behind[:n_ocn]) and not np.array_equal(at[:n_ocn], prev[:n_ocn])
    assert np.array_equal(at[n_ocn:], prev[n_ocn:]) and not np.array_equal(at[n_ocn:], behind[n_ocn:])
    assert at[0] > 0.0


def test_res_files_without_surface_columns(built, tmp_path):
    """ctrl_data_save_sig_ocn_sur = .FALSE.: two-column T / S files, (mol, mol kg-1) for bulk tracers, (mol, o/oo) for isotopes
    (biogem_data_ascii.f90:59-71, 737-747, 765-776, 797-802)."""
    out = tmp_path / "nosur"
    write_series(str(out), with_sur=False)
    assert open(out / "biogem_series_ocn_temp.res").read() == " % time (yr) / temperature (degrees C)\n"
    assert open(out / "biogem_series_ocn_ALK.res").read() == " % time (yr) / global ALK (mol) / global ALK (mol kg-1)\n"
    assert open(out / "biogem_series_ocn_DIC_14C.res").read() == " % time (yr) / global DIC_14C (mol) / global DIC_14C (o/oo)\n"
    sig = np.zeros(3 + 3 * L + LA)
    sig[0], sig[1] = 0.25, 0.25 * 1.0e21
    sig[3], sig[4], sig[5] = 0.25 * 273.15, 0.25 * 35.0, 0.25 * 2.0e-3
    sig[3 + 4] = 0.25 * 2.0e-3 * 1.176e-12 / (1.0 + 1.176e-12)          # 14C at the standard ratio: delta = 0
    write_series(str(out), sig, t_yr=7.125, with_sur=False)
    assert open(out / "biogem_series_ocn_temp.res").read().split("\n")[1] == "       7.125    0.000000"
    assert open(out / "biogem_series_ocn_sal.res").read().split("\n")[1] == "       7.125   35.000000"
    assert open(out / "biogem_series_ocn_DIC.res").read().split("\n")[1] == "       7.125  0.2000000E+19  0.2000000E-02"
    l14 = open(out / "biogem_series_ocn_DIC_14C.res").read().split("\n")[1]
    assert l14[:27] == "       7.125  0.2352000E+07" and l14[27:] in ("       0.000", "      -0.000")


def test_extended_series_files(built, tmp_path):
    """fexport_*, fseaair_*, focnatm_*, misc_* (cg_biogem_series_write_ext) in the reference's headers and edit descriptors
    (biogem_data_ascii.f90:107-197, 320-400; 955-1096, 1245-1340) from hand-set integrals."""
    from cgenie_b200.series import write_series_ext
    nL, nS, nA = 16, 9, 8
    sig = np.zeros(3 + 3 * nL + nA)
    sig2 = np.zeros(8 + nS + 2 * nA)
    t = 0.5
    sig[0] = t
    atm = sig[3 + 3 * nL:]
    atm[2] = t * 278.0e-6
    atm[3] = t * 278.0e-6 * 0.011057        # r13C
    atm[4] = t * 278.0e-6 * 1.2e-12
    sig2[0:8] = t * np.array([1.5e13, 1.25, 2.0e13, -0.0123, 0.0234, -0.002, 0.0101, 3.25])
    fe = sig2[8:8 + nS]
    fe[0] = t * 8.0e14; fe[1] = t * 8.0e14 * 0.0109; fe[3] = t * 7.5e12; fe[4] = t * 1.0e14
    oa = sig2[8 + nS:8 + nS + nA]
    oa[2] = t * -2.5e13; oa[3] = t * -2.5e13 * 0.0111; oa[5] = t * 4.0e12
    sa = sig2[8 + nS + nA:]
    sa[2] = t * 1.0e12; sa[5] = t * -3.0e12
    out = tmp_path / "o"
    write_series_ext(str(out), ocn_tot_A=3.6e14)
    files = sorted(p.name for p in out.iterdir())
    assert "biogem_series_fexport_POC_frac2.res" not in files and "biogem_series_fseaair_pCO2.res" in files
    assert len(files) == 7 + 6 + 6 + 4
    assert open(out / "biogem_series_fexport_POC.res").read() == " % time (yr) / global POC flux (mol yr-1) / global POC density (mol m-2 yr-1)\n"
    assert open(out / "biogem_series_misc_opsi.res").read().startswith(" % time (yr) / global min overturning (Sv) / global max overturning (Sv) / Atlantic min")
    write_series_ext(str(out), sig, sig2, t_yr=12.5, ocn_tot_A=3.6e14)
    line = lambda n: open(out / ("biogem_series_%s.res" % n)).read().split("\n")[1]
    assert line("fexport_POC") == "      12.500  0.8000000E+15  0.2222222E+01"
    assert line("fexport_POP") == "      12.500  0.7500000E+13  0.2083333E-01"
    got = line("fexport_POC_13C")
    r = 0.0109 / (1 - 0.0109)
    assert got[:27] == "      12.500  0.8720000E+13" and abs(float(got[27:]) - 1000.0 * (r / 0.011202 - 1.0)) < 1e-3 and len(got) == 12 + 15 + 14
    assert line("focnatm_pCO2") == "      12.500 -0.2500000E+14      -0.069"
    assert line("focnatm_pO2") == "      12.500  0.4000000E+13       0.011"
    assert line("fseaair_pO2") == "      12.500 -0.3000000E+13      -0.008"
    assert line("misc_seaice") == "      12.500  0.1500E+14    4.167  0.2000E+14    1.250"
    assert line("misc_opsi") == "      12.500  -19.588   37.264   -3.185   16.084"     # 1592.5 * 0.0234 = 37.26449999999999...
    assert line("misc_SLT") == "      12.500    3.250000"
    d13 = 1000.0 * (0.011057 / (1 - 0.011057) / 0.011202 - 1.0)
    d14 = 1000.0 * (1.2e-12 / (1 - 1.2e-12) / 1.176e-12 - 1.0)
    D14 = 1000.0 * ((1.0 + d14 / 1000.0) * 0.975 ** 2 / (1.0 + d13 / 1000.0) ** 2 - 1.0)
    assert abs(float(line("misc_atm_D14C").split()[1]) - D14) < 1e-3
    # a topography without the Atlantic columns prints two overturning values
    write_series_ext(str(tmp_path / "p"), atlantic=False)
    write_series_ext(str(tmp_path / "p"), sig, sig2, t_yr=12.5, atlantic=False)
    assert open(tmp_path / "p" / "biogem_series_misc_opsi.res").read().split("\n")[1] == "      12.500  -19.588   37.264"


def test_series_saver_extended(built, tmp_path):
    """SeriesSaver(extended=True): switches the extended integrals on before the first window, and at every save writes the export /
    air-sea flux / misc files next to the ocn_* / atm_* ones from "bg_sig" and "bg_sig2" (recording engine, constant integrands)."""
    from cgenie_b200.series import SeriesSaver

    class Eng(_FakeEngine):
        def __init__(self):
            super().__init__()
            self.extended_calls = 0

        def biogem_sig_extended(self):
            assert self.updates == 0            # before the first BIOGEM step of interest
            self.extended_calls += 1

        def const(self, name):
            assert name == "bg_ocn_tot_A"
            return np.array([3.6e14])

        def get(self, name, member):
            if name == "bg_sig":
                return super().get(name, member)
            assert name == "bg_sig2"
            x = np.zeros(8 + 9 + 2 * LA)
            x[0], x[3], x[4] = self.t * 1.0e13, self.t * -0.01, self.t * 0.02
            x[8] = 48 * self.t * 1.0e13             # export: summed per step, not time weighted
            x[8 + 9 + 2] = self.t * -2.0e13
            return x

    nyear, kb = 96, 10
    genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / nyear
    tick = int(round(1000.0 * genie_timestep))
    dts = float(kb) * genie_timestep
    e = Eng()
    s = SeriesSaver(e, tmp_path / "x", t_runtime=2.0, t_start=0.0, sig_dt=1.0, extended=True, world="worjh2")
    assert e.extended_calls == 1
    for k in range(kb, 2 * 5 * nyear + 1, kb):
        s.step(dts, k * tick)
    assert s.saved == [0.5, 1.5]
    opsi = open(tmp_path / "x" / "biogem_series_misc_opsi.res").read().split("\n")
    assert len(opsi) == 4 and opsi[1] == "       0.500  -15.925   31.850    0.000    0.000" and opsi[2][:12] == "       1.500"
    assert open(tmp_path / "x" / "biogem_series_fexport_POC.res").read().split("\n")[2] == "       1.500  0.4800000E+15  0.1333333E+01"
    assert open(tmp_path / "x" / "biogem_series_focnatm_pCO2.res").read().split("\n")[1] == "       0.500 -0.2000000E+14      -0.056"
    assert len(open(tmp_path / "x" / "biogem_series_ocn_temp.res").read().split("\n")) == 4
    # a topography without the Atlantic columns
    e2 = Eng()
    s2 = SeriesSaver(e2, tmp_path / "y", t_runtime=1.0, sig_dt=1.0, extended=True, world="p0055c")
    for k in range(kb, 5 * nyear + 1, kb):
        s2.step(dts, k * tick)
    assert open(tmp_path / "y" / "biogem_series_misc_opsi.res").read().split("\n")[1] == "       0.500  -15.925   31.850"


def test_frozen_tracer_tables_are_the_references():
    """The frozen 16 / 8 / 9 tracer selection as the Python host carries it (restart.OCN_TRACERS / ATM_TRACERS / SED_TRACERS with their
    global indices, series.*_TYPE / *_DEP) against data/main/tracer_define.{ocn,atm,sed} of the reference (tests/golden/
    ref_tracer_define.json): names, long names, types, and the dependencies mapped to compact indices."""
    import json
    import os
    from cgenie_b200 import restart, series
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tracer_define.json")))["tracers"]
    by = {k: {r["index"]: r for r in rows} for k, rows in ref.items()}
    sel = {"ocn": (restart.OCN_IDS, [n for n, _ in restart.OCN_TRACERS], [l for _, l in restart.OCN_TRACERS], series.OCN_TYPE, series.OCN_DEP),
           "atm": ([i for i, _, _ in restart.ATM_TRACERS], [n for _, n, _ in restart.ATM_TRACERS], [l for _, _, l in restart.ATM_TRACERS],
                   series.ATM_TYPE, series.ATM_DEP),
           "sed": (restart.SED_IDS, [n for n, _ in restart.SED_TRACERS], [l for _, l in restart.SED_TRACERS], series.SED_TYPE, series.SED_DEP)}
    for kind, (ids, names, longs, types, deps) in sel.items():
        assert len(ids) == len(names) == len(types) == len(deps)
        for q, gi in enumerate(ids):
            r = by[kind][gi]
            assert r["name"] == names[q] and r["long_name"] == longs[q], (kind, gi, r, names[q], longs[q])
            assert r["type"] == types[q], (kind, r["name"], r["type"], types[q])
            assert ids.index(r["dep"]) == deps[q], (kind, r["name"], r["dep"], deps[q])
