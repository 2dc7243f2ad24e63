"""SEDGEM / ROKGEM coupler calls on the device-resident interface arrays (cg_cpl_flux_ocnsed, cg_cpl_comp_ocnsed,
cg_reinit_flux_rokocn) -- run with -m gpu on a B200.

The accumulations are one multiply-add (or multiply-add-divide) per element without contraction, so the device sums must equal
bit for bit what numpy makes of the device's own sfxsed1 / sfcocn1; against the oracle the sums carry the tolerance of the
fields they integrate (1e-10 relative per step, BASELINE.json north_star)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_biogem import CFG, OKW, I, J, L, LS

pytestmark = pytest.mark.gpu


def test_sedgem_coupler_sums_on_device(built, tmp_path):
    materialise(str(tmp_path), CFG)
    M = 3
    pert = {"par_bio_k0_PO4": np.array([1.9e-6, 1.7e-6, 2.3e-6]), "par_bio_red_POC_CaCO3": np.array([0.044, 0.03, 0.06])}
    o = Oracle(**OKW)
    o.biogem_setup(par_bio_k0_PO4=1.9e-6, par_bio_red_POC_CaCO3=0.044)
    with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
        e.set_tracer_variant("strict")
        genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
        tick = int(round(1000.0 * genie_timestep))
        dts = float(2 * 5) * genie_timestep
        flux = [np.zeros(J * I * LS) for _ in range(M)]
        comp = [np.zeros(J * I * L) for _ in range(M)]
        for m in range(M):
            e.put("sfxsumrok1", np.full(J * I * L, 2.5), m)
        for blk in range(1, 5):
            if blk <= 2:          # whole iterations on the device, coupler calls behind cg_run
                e.run(10)
            else:                 # module by module, coupler calls where genie.f90 makes them (after step_biogem)
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
            ocnstep = 10 * blk // 5
            e.cpl_flux_ocnsed(dts)
            e.cpl_comp_ocnsed(ocnstep, 2, 8)
            e.reinit_flux_rokocn()
            if blk > 2:
                e.biogem_tracercoupling()
                e.biogem_climate()
                e.atchem_step(dts)
            o.run(10)
            o.L.cgo_cpl_flux_ocnsed(o.h, dts)
            o.L.cgo_cpl_comp_ocnsed(o.h, ocnstep, 2, 8)
            w = ((ocnstep - 2) % 8) // 2
            for m in range(M):
                flux[m] = flux[m] + dts * e.get("sfxsed1", m)
                comp[m] = (float(w) * comp[m] + e.get("sfcocn1", m)) / float(w + 1)
                assert np.array_equal(e.get("sfxsumsed", m), flux[m]), (blk, m)
                assert np.array_equal(e.get("sfcsumocn", m), comp[m]), (blk, m)
                assert np.all(e.get("sfxsumrok1", m) == 0.0)
            for name in ("sfxsumsed", "sfcsumocn"):
                d, r = e.get(name, 0), o.f(name)
                scale = np.abs(r).reshape(J * I, -1).max(axis=0)[None, :] + 1e-300
                assert (np.abs(d - r).reshape(J * I, -1) / scale).max() <= 1e-9, (name, blk)
        assert np.abs(flux[0]).max() > 0.0 and not np.array_equal(flux[0], flux[1])     # members differ, rain is non-zero
        assert int(e.health().sum()) == 0
