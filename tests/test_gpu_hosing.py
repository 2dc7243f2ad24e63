"""Row a10 of SURVEY 8: get_hosing (goldstein.f90:3086-3129) switched ON -- constant hosing plus a trend, a hosing period that
ends inside the run, and per-member hosing amplitudes -- through cg_run against the oracle.  Run with -m gpu on a B200.

The freshwater hosing enters the salinity surface flux of the hosing region (goldstein.f90:149-170); the device path is
k_hosing + k_gold_pre.  Strict variant: everything but surflux's libm calls is bit-identical, so the bar is the per-step one."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_parity import interior

pytestmark = pytest.mark.gpu


def test_hosing_on_matches_oracle(built, tmp_path):
    I = J = 36
    K, L = 8, 2
    hos = np.array([0.2, 0.05, 0.4])
    # nyears_hosing = 1 -> the hosing acts for the first 100 ocean steps.  Both sides start from the initial state with the ocean
    # step counter at 94, so the run sees 6 steps with hosing (and its trend) and 14 steps after the period has ended
    first, nsteps = 94, 20
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_hosing": 0.2, "go_hosing_trend": 50.0, "go_nyears_hosing": 1})
    with Ensemble(str(job), n_members=3, perturb={"hosing": hos}) as e:
        e.set_tracer_variant("strict")
        e.set_koverall(5 * first)
        e.run(5 * nsteps)
        got = [{n: e.get(n, m) for n in ("ts", "rho")} for m in range(3)]
        assert int(e.health().sum()) == 0
    worst = 0.0
    for m in range(3):
        o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, hosing=float(hos[m]), hosing_trend=50.0, nyears_hosing=1)
        o.set("istep_ocn", first)
        o.run(5 * 3)
        assert o.f("fw_hosing").max() > 0.0                       # the hosing is acting
        o.run(5 * (nsteps - 3))
        assert np.all(o.f("fw_hosing") == 0.0)                    # ... and has stopped
        ts = interior(o, "ts")
        scale = np.abs(ts.reshape(-1, L)).max(axis=0)
        err = np.abs(got[m]["ts"] - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-10 * nsteps, (m, float(err.max()))
    o0 = Oracle("worbe2", maxk=K, maxl=L, nyear=100)
    o0.set("istep_ocn", first)
    o0.run(5 * nsteps)
    assert np.abs(interior(o0, "ts") - got[0]["ts"]).max() > 1e-7    # a run without hosing differs
    print("hosing on, 3 members, %d ocean steps: worst per-cell relative difference %.2e" % (nsteps, worst))
