"""The water-column sweep of step_biogem in its packets / cells form (k_bg_step PART 3 + k_bg_cell: the per-cell half of the step
fused with biogem_tracercoupling's update, bio_part's rescaling left pending for the next reader) against the one-kernel sweep +
k_tc_apply (CG_BG_PD=0) -- run with -m gpu on a B200.  Same expressions in the same order: every field must be BIT-IDENTICAL, through
cg_run's pipelined schedule and through the per-module calls, with diagnostics reading bio_part in between."""
import os

import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.sharding import perturbation_table
from test_gpu_biogem import CFG

pytestmark = pytest.mark.gpu
NAMES = ("ts", "ocn", "bio_part", "bg_M", "bg_rM", "atm", "carbH", "settle_k1", "sfcocn1", "sfxsed1", "tq", "u")


def run(jobdir, M, tab, pd, per_module):
    os.environ["CG_BG_PD"] = "1" if pd else "0"
    try:
        with Ensemble(jobdir, n_members=M, perturb=tab) as e:
            e.set_tracer_variant("col")
            e.run(5 * 20)                                   # 10 blocks through cg_run
            snap1 = {n: e.get(n, M - 1) for n in ("bio_part", "ocn")}      # a host read between blocks (applies a pending rescaling)
            if per_module:
                genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
                tick = int(round(1000.0 * genie_timestep))
                dts = float(2 * 5) * genie_timestep
                for k in range(101, 141):
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
                        if k % 20 == 0:
                            e.biogem_slice_update(dts)      # integrates bio_part
                        e.atchem_step(dts)
                e.set_koverall(140)
            else:
                e.run(5 * 8)
            e.run(5 * 6)
            assert int(e.health().sum()) == 0
            out = {n: np.stack([e.get(n, m) for m in (0, M // 2, M - 1)]) for n in NAMES}
            out["snap_part"], out["snap_ocn"] = snap1["bio_part"], snap1["ocn"]
            if per_module:
                out["sl_part"] = e.get("sl_part", 1)
            return out
    finally:
        os.environ.pop("CG_BG_PD", None)


@pytest.mark.parametrize("per_module", [False, True])
def test_packets_cells_form_is_bit_identical(built, tmp_path, per_module):
    materialise(str(tmp_path), CFG)
    M = 40
    tab = perturbation_table(M, biogem=True)
    a = run(str(tmp_path), M, tab, True, per_module)
    b = run(str(tmp_path), M, tab, False, per_module)
    for n in a:
        assert np.array_equal(a[n], b[n]), (per_module, n, float(np.abs(a[n] - b[n]).max()))
    assert np.abs(a["bio_part"]).max() > 1e-8
