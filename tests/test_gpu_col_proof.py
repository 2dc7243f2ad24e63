"""The benched tracer step ('col') proven per cell, for all 128 lanes, along the spin-up and at the state the bench times
(VERDICT r1 "what's weak" 2; SURVEY section 7, hard part 2: report the flip rate) -- run with -m gpu on a B200.

tests/col_check.py takes ONE tstepo in 'strict' (bit-identical to the oracle, tests/test_gpu_parity.py) and one in 'col' from
the same state: every cell outside columns whose convection decisions differ must agree to 1e-10 relative to the cell's own
value; flipped columns are counted (flip rate) and must conserve every tracer's column inventory."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.sharding import perturbation_table
from col_check import check_step
from test_gpu_biogem import CFG

pytestmark = pytest.mark.gpu


def test_col_per_cell_all_lanes_along_the_spin_up(built, tmp_path):
    """128 perturbed members (the bench's table) from the uniform initial state -- the neutrally stable regime where
    trajectories of different arithmetic part -- through the first 12 ocean steps one by one, then at 1, 5, 25 and 100 model
    years (the bench's state: 100 years of spin-up)."""
    materialise(str(tmp_path), CFG)
    M = 128
    tab = perturbation_table(M, biogem=True)
    report = []
    with Ensemble(str(tmp_path), n_members=M, perturb=tab) as e:
        e.set_tracer_variant("col")
        done = 0
        for at in list(range(1, 13)) + [96, 5 * 96, 25 * 96, 100 * 96]:
            e.set_tracer_variant("col")
            e.run(5 * (at - done))
            done = at
            st = check_step(e, tol=1e-10, what="ocean step %d:" % at, strict_assert=False)
            report.append((at, st))
        assert int(e.health().sum()) == 0
    for at, st in report:      # every checkpoint is held to the bar; the loop above only collects, so that a failure shows all of them
        assert st["worst_unflipped"] <= 1e-10, (at, st)
        assert st.get("worst_flipped_inventory", 0.0) <= 1e-12, (at, st)
    rates = [st["flip_rate"] for _, st in report]
    print("flip rate per ocean step along the spin-up:", ["%d: %.1e" % (at, r) for (at, _), r in zip(report, rates)])
    # the flips are a property of the first, neutrally stable months; a stratified ocean has (next to) none
    assert report[-1][1]["flip_rate"] <= 1e-3 and max(rates) <= 0.5      # step 1 starts from the uniform, everywhere neutral state
    assert report[-1][1]["convecting_columns"] > 0
