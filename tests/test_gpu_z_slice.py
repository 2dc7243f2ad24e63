"""Time-slice diagnostics of BIOGEM on the device (cg_biogem_slice_update, k_bg_slice; SURVEY 8f row 1, time-slice part) against
the oracle's restatement of diag_biogem_timeslice (biogem.f90:2421-2699) -- run with -m gpu on a B200.

What is compared: the 3-D carbonate re-solve of every wet cell ([H+] kept per cell from sub_init_carb on) and the window
integrals int_ocn / int_bio_part / int_carb / int_carbconst / int_carbisor / int_t _timeslice after three BIOGEM steps inside a
save window, per CELL, relative to the cell's own value.  Bar: 1e-10 (BASELINE.json north_star); the tracer integrals are
transcendental-free but carry the 1e-14 the tracers themselves are apart after 4 BIOGEM steps."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_biogem import CFG, OKW, I, J, K, L, LS, wet_masks

pytestmark = pytest.mark.gpu

N_IC, N_CC, N_ICI = 10, 17, 8


def test_slice_integrals_match_oracle(built, tmp_path):
    materialise(str(tmp_path / "job"), CFG)
    M = 2
    pert = {"par_bio_k0_PO4": np.array([1.9e-6, 2.4e-6])}
    o = Oracle(**OKW)
    o.biogem_setup(par_bio_k0_PO4=1.9e-6)
    with Ensemble(str(tmp_path / "job"), n_members=M, perturb=pert) as e:
        e.set_tracer_variant("strict")
        genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
        dts = float(2 * 5) * genie_timestep
        _, wet3 = wet_masks(o)
        wet = wet3.ravel()
        # sub_init_carb below the surface: [H+] of every wet cell from the common seed
        H0d, H0o = e.get("carbH3", 0), o.f("carb3").reshape(-1, N_IC)[:, 0]
        below = wet.copy()
        below[(K - 1) * I * J:] = False
        assert np.max(np.abs(H0d[below] - H0o[below]) / H0o[below]) <= 1e-12
        e.run(40)
        o.run(40)
        e.biogem_slice_reset()
        for blk in range(3):
            e.run(10)
            o.run(10)
            e.biogem_slice_update(dts)
            o.L.cgo_biogem_slice_update(o.h)
        t_d, t_o = e.get("sl_t", 0), o.f("sl_t")
        assert t_d[0] == t_o[0] and abs(t_d[0] - 3 * dts / (3600.0 * 24.0 * 365.25)) < 1e-15
        worst = {}
        for name, n in (("sl_ocn", L), ("sl_part", LS), ("sl_carb", N_IC), ("sl_carbconst", N_CC), ("sl_carbisor", N_ICI)):
            d = e.get(name, 0).reshape(-1, n)[wet]
            r = o.f(name).reshape(-1, n)[wet]
            floor = np.maximum(1e-3 * np.abs(r).max(axis=0), 1e-300)
            worst[name] = float(np.max(np.abs(d - r) / np.maximum(np.abs(r), floor)))
        print("time-slice integrals, worst per-cell relative difference:", {k: "%.2e" % v for k, v in worst.items()})
        assert worst["sl_ocn"] <= 1e-12 and worst["sl_part"] <= 1e-10, worst
        for k in ("sl_carb", "sl_carbconst", "sl_carbisor"):
            assert worst[k] <= 1e-10, worst
        # the surface cell's [H+] is the seed of step_biogem's next solve: the diagnostic must leave it as the oracle's does
        Hs_d, Hs_o = e.get("carbH", 0), o.f("carb").reshape(J * I, -1)[:, 0]
        wet2 = wet3[K - 1].ravel()
        assert np.max(np.abs(Hs_d[wet2] - Hs_o[wet2]) / Hs_o[wet2]) <= 1e-10
        # another uptake rate, another deep DIC integral; same window length
        assert e.get("sl_t", 1)[0] == t_d[0]
        assert np.any(e.get("sl_ocn", 1) != e.get("sl_ocn", 0))
        # the run goes on as the oracle's does (the feedback through the surface seed included)
        e.run(10)
        o.run(10)
        od, oo = e.get("ocn", 0).reshape(-1, L)[wet], o.f("ocn").reshape(-1, L)[wet]
        floor = np.maximum(1e-3 * np.abs(oo).max(axis=0), 1e-300)
        assert float(np.max(np.abs(od - oo) / np.maximum(np.abs(oo), floor))) <= 1e-10
        e.biogem_slice_reset()
        assert np.all(e.get("sl_ocn", 1) == 0.0) and e.get("sl_t", 0)[0] == 0.0
        assert int(e.health().sum()) == 0
