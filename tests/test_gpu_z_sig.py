"""Window integrals of BIOGEM's time series on the device (cg_biogem_sig_update, k_bg_sig_sums / k_bg_sig_acc; SURVEY 8f row 1,
time-series part) against the oracle's restatement of diag_biogem_timeseries (biogem.f90:2836-2917) -- run with -m gpu on a B200.
Bar: 1e-10 relative (BASELINE.json north_star); the .res files written from the device integrals equal those written from the
oracle's character for character wherever the printed digits are not at a rounding boundary (checked as numbers to 1e-6)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.series import write_series
from oracle_lib import Oracle
from test_gpu_biogem import CFG, OKW, L, LA

pytestmark = pytest.mark.gpu


def test_sig_integrals_match_oracle(built, tmp_path):
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
        e.run(40)
        o.run(40)
        assert np.all(e.get("bg_sig", 0) == 0.0)
        for blk in range(5, 9):
            if blk <= 6:          # whole iterations on the device, the diagnostic behind cg_run
                e.run(10)
            else:                 # module by module; the diagnostic behind the whole BIOGEM / ATCHEM block, where the oracle takes it
                for k in range(10 * (blk - 1) + 1, 10 * blk + 1):
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
                        e.atchem_step(dts)
            e.biogem_sig_update(dts, 1000.0)
            o.run(10)
            o.L.cgo_biogem_sig_update(o.h, 1000.0)
        d, r = e.get("bg_sig", 0), o.f("bg_sig")
        assert d.size == 3 + 3 * L + LA and d[0] == r[0] and abs(d[0] - 4 * dts / (3600.0 * 24.0 * 365.25)) < 1e-15
        rel = np.abs(d - r) / np.maximum(np.abs(r), 1e-300)
        print("bg_sig worst relative difference %.2e at %d" % (rel.max(), int(rel.argmax())))
        assert rel.max() <= 1e-10, (rel.max(), int(rel.argmax()))
        d1 = e.get("bg_sig", 1)
        assert d1[0] == d[0] and d1[3 + L + 5] != d[3 + L + 5]          # another uptake rate, another mean surface PO4
        for who, sig in (("dev", d), ("ora", r)):
            write_series(str(tmp_path / who), None)
            write_series(str(tmp_path / who), sig, t_yr=0.146)
        for n in ("ocn_temp", "ocn_DIC", "ocn_DIC_13C", "ocn_PO4", "atm_pCO2", "atm_pCO2_14C"):
            a = open(tmp_path / "dev" / ("biogem_series_%s.res" % n)).read().split("\n")
            b = open(tmp_path / "ora" / ("biogem_series_%s.res" % n)).read().split("\n")
            assert a[0] == b[0] and len(a) == len(b) == 3
            va, vb = np.array(a[1].split(), dtype=float), np.array(b[1].split(), dtype=float)
            assert np.allclose(va, vb, rtol=1e-6, atol=2e-3), n
        e.biogem_sig_reset()
        assert np.all(e.get("bg_sig", 2) == 0.0)
        assert int(e.health().sum()) == 0
