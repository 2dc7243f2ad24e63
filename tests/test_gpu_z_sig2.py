"""The flux, export and "misc" integrals of BIOGEM's time series on the device (cg_biogem_sig_extended: k_bg_settle_sur,
k_bg_sig2_stage / _sums / _acc) against the oracle's restatement of diag_biogem_timeseries (biogem.f90:2870-2883,
2926-2964, 3058-3062; sub_calc_psi biogem_box.f90:3796-3852) at genie.f90's own call point (:401-405: behind biogem_climate, ahead
of cpl_flux_ocnatm and ATCHEM) -- run with -m gpu on a B200.  Bar: 1e-10 relative per integral; the overturning extrema are taken
in the reference's summation order (bit-exact against the oracle for the strict variant's velocities)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.series import write_series_ext
from oracle_lib import Oracle
from test_gpu_biogem import CFG, OKW, L, LA

pytestmark = pytest.mark.gpu
LS = 9


def test_extended_sig_integrals_match_oracle(built, tmp_path):
    materialise(str(tmp_path / "job"), CFG)
    M = 3
    pert = {"par_bio_k0_PO4": np.array([1.9e-6, 1.7e-6, 2.3e-6])}
    o = Oracle(**OKW)
    o.biogem_setup(par_bio_k0_PO4=1.9e-6)
    with Ensemble(str(tmp_path / "job"), n_members=M, perturb=pert) as e:
        e.set_tracer_variant("strict")
        genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
        tick = int(round(1000.0 * genie_timestep))
        dts = float(2 * 5) * genie_timestep
        e.run(1400)                   # towards the end of the third model year: the first sea ice, an overturning cell, export
        o.run(1400)
        e.biogem_sig_extended()
        e.biogem_sig_reset()
        o.L.cgo_biogem_sig_auto(o.h, 1, 1000.0)
        assert np.all(e.get("bg_sig2", 0) == 0.0)
        nblk = 6
        for k in range(1401, 1400 + 10 * nblk + 1):
            if k % 5 == 1:
                e.surflux()
            e.step_embm()
            if k % 5 == 0:
                e.step_seaice()
                e.step_goldstein()
            if k % 10 == 0:
                e.biogem_forcing(k * tick)
                e.biogem_step(dts, k * tick)
                e.biogem_tracercoupling()
                e.biogem_climate()
                e.biogem_sig_update(dts, 1000.0)
                e.atchem_step(dts)
        o.run(10 * nblk)
        d, r = e.get("bg_sig2", 0), o.f("bg_sig2")
        s, rs = e.get("bg_sig", 0), o.f("bg_sig")
        assert d.size == 8 + LS + 2 * LA and s[0] == rs[0]
        names = (["seaice", "seaice_th", "seaice_vol", "opsi_min", "opsi_max", "opsia_min", "opsia_max", "SLT"] +
                 ["fexport%d" % q for q in range(LS)] + ["focnatm%d" % q for q in range(LA)] + ["airsea%d" % q for q in range(LA)])
        scale = np.maximum(np.abs(r), 1e-300)
        rel = np.abs(d - r) / scale
        for n, a, b, x in zip(names, d, r, rel):
            print("%-10s device % .15e oracle % .15e rel %.1e" % (n, a, b, x))
        assert rel.max() <= 1e-10, (names[int(rel.argmax())], float(rel.max()))
        # something to compare: ice, an overturning cell, export of POC and CaCO3, CO2 and O2 exchange, net flux incl. the restoring
        assert r[0] > 0 and r[2] > 0 and r[3] < 0 < r[4] and r[5] < 0 < r[6]
        assert r[8] > 0 and r[8 + 4] > 0 and r[8 + LS + 2] != 0 and r[8 + LS + LA + 5] != 0
        assert np.all(d[8 + LS:8 + LS + 2] == 0) and np.all(d[8 + LS + LA:8 + LS + LA + 2] == 0)
        d1 = e.get("bg_sig2", 2)
        assert d1[8] != d[8]                                   # another uptake rate, another export
        A = float(e.const("bg_ocn_tot_A")[0])
        for who, (sg, sg2) in (("dev", (s, d)), ("ora", (rs, r))):
            write_series_ext(str(tmp_path / who), ocn_tot_A=A)
            write_series_ext(str(tmp_path / who), sg, sg2, t_yr=2.979, ocn_tot_A=A)
        for n in ("fexport_POC", "fexport_POC_13C", "fexport_CaCO3", "fseaair_pCO2", "fseaair_pCO2_13C", "focnatm_pCO2", "focnatm_pO2",
                  "misc_seaice", "misc_opsi", "misc_atm_D14C", "misc_SLT"):
            a = open(tmp_path / "dev" / ("biogem_series_%s.res" % n)).read().split("\n")
            b = open(tmp_path / "ora" / ("biogem_series_%s.res" % n)).read().split("\n")
            assert a[0] == b[0] and len(a) == len(b) == 3, n
            va, vb = np.array(a[1].split(), dtype=float), np.array(b[1].split(), dtype=float)
            assert np.allclose(va, vb, rtol=1e-6, atol=2e-3), n
        e.biogem_sig_reset()
        assert np.all(e.get("bg_sig2", 1) == 0.0)
        assert int(e.health().sum()) == 0
