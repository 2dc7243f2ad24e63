"""SURVEY 8f row 4, second option: the thermobaric equation of state (ieos = 1: eos / eosd goldstein.f90:3048-3082, the
re-evaluations inside co :2692-2730, the vertically local rho of tstepo :2396-2408) through cg_run against the oracle -- run with
-m gpu on a B200.  'strict' tracer variant: one tstepo from a common state is BIT-EXACT (ts, rho, cost); a run from the initial
state is held to the per-step bar (surflux's libm calls are the only difference between device and oracle)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_parity import inject, interior

pytestmark = pytest.mark.gpu
I = J = 36
K, L = 8, 2


def test_thermobaric_tracer_step_bit_exact(built, tmp_path):
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_ieos": 1})
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, ieos=1)
    o.run(5 * 150)
    with Ensemble(str(job), n_members=1) as e:
        e.set_tracer_variant("strict")
        inject(e, o, 0)
        e.set_koverall(5 * 150)
        for step in range(3):
            e._ck(e.L.cg_tracer_step(e.h, 1))
            o.call("tstepo")
            assert np.array_equal(e.get("ts", 0), interior(o, "ts")), step
            assert np.array_equal(e.get("rho", 0), interior(o, "rho")), step
            assert np.array_equal(e.get("cost", 0), o.f("cost")), step
        assert o.f("cost").sum() > 0


def test_thermobaric_run_matches_oracle(built, tmp_path):
    nsteps = 40
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_ieos": 1})
    scf = np.array([2.0, 1.7])
    with Ensemble(str(job), n_members=2, perturb={"scf": scf}) as e:
        e.set_tracer_variant("strict")
        e.run(5 * nsteps)
        got = [{n: e.get(n, m) for n in ("ts", "rho", "u")} for m in range(2)]
        e.set_tracer_variant("col")                    # no thermobaric term in the column kernel: the generic kernels run
        assert e.tracer_variant_active() == "fast"
        e.run(5)
        assert int(e.health().sum()) == 0
    worst = 0.0
    for m in range(2):
        o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, ieos=1, scf=float(scf[m]))
        o.run(5 * nsteps)
        ts = interior(o, "ts")
        scale = np.abs(ts.reshape(-1, L)).max(axis=0)
        err = np.abs(got[m]["ts"] - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-10 * nsteps, (m, float(err.max()))
        rho = interior(o, "rho")
        assert np.abs(got[m]["rho"] - rho).max() <= 1e-10 * nsteps * np.abs(rho).max()
    o0 = Oracle("worbe2", maxk=K, maxl=L, nyear=100)
    o0.run(5 * nsteps)
    assert np.abs(interior(o0, "ts") - got[0]["ts"]).max() > 1e-6     # the option is acting
    print("ieos=1: worst per-cell relative difference after %d ocean steps %.2e" % (nsteps, worst))


@pytest.mark.parametrize("ieos", [0, 1])
def test_mueller_convection_bit_exact(built, tmp_path, ieos):
    """iconv = 1 (coshuffle, goldstein.f90:2781-2841, + the depth diagnostic :2766-2770), alone and with the thermobaric term: three
    tstepo from a common spun-up state, strict variant, bit for bit (ts, rho and the diagnostic cost = dsc * zw(maxk-1-icosd))."""
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_iconv": 1, "go_ieos": ieos})
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, iconv=1, ieos=ieos)
    o.run(5 * 120)
    with Ensemble(str(job), n_members=1) as e:
        e.set_tracer_variant("strict")
        inject(e, o, 0)
        e.set_koverall(5 * 120)
        for step in range(3):
            e._ck(e.L.cg_tracer_step(e.h, 1))
            o.call("tstepo")
            assert np.array_equal(e.get("ts", 0), interior(o, "ts")), step
            assert np.array_equal(e.get("rho", 0), interior(o, "rho")), step
            assert np.array_equal(e.get("cost", 0), o.f("cost")), step
        assert o.f("cost").min() < 0.0                      # a depth (m, negative), not a count
        assert int(e.health().sum()) == 0


@pytest.mark.parametrize("ieos", [0, 1])
def test_mueller_convection_run_matches_oracle(built, tmp_path, ieos):
    """... and 40 ocean steps of the whole model from the initial state with the scheme on, per-step bar."""
    nsteps = 40
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_iconv": 1, "go_ieos": ieos})
    with Ensemble(str(job), n_members=1) as e:
        e.set_tracer_variant("strict")
        e.run(5 * nsteps)
        got = {n: e.get(n, 0) for n in ("ts", "cost")}
        assert int(e.health().sum()) == 0
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, iconv=1, ieos=ieos)
    o.run(5 * nsteps)
    ts = interior(o, "ts")
    scale = np.abs(ts.reshape(-1, L)).max(axis=0)
    err = np.abs(got["ts"] - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
    print("iconv=1 ieos=%d: worst per-cell relative difference after %d ocean steps %.2e" % (ieos, nsteps, float(err.max())))
    assert err.max() <= 1e-10 * nsteps, float(err.max())
    assert np.array_equal(got["cost"], o.f("cost"))
