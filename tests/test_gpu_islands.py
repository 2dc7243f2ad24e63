"""Topographies with more than one island (the reference's data/goldstein worlds p0055c: two islands, p0251a: three): the
barotropic closure with one path integral per island and matmult (goldstein.f90:203-230, 3470-3492; island
goldstein_lib.f90:186-241) on the device against the oracle -- run with -m gpu on a B200.

The momentum step has no libm call: psi, ub and u after the first ocean step are BIT-EXACT in the strict variant (members with
their own drag, i.e. their own unit island solves and erisl matrix); a run of 30 ocean steps is held to the per-step bar; the
production barotropic solve (blocked substitution) agrees with the strict one to rounding."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_parity import interior

pytestmark = pytest.mark.gpu
I = J = 36
K, L = 16, 2


@pytest.mark.parametrize("world,nisl", [("p0055c", 2), ("p0251a", 3)])
def test_multi_island_closure_matches_oracle(built, tmp_path, world, nisl):
    cfg = "eb_go_gs_%s_36x36x16" % world
    materialise(str(tmp_path), cfg)
    adrag = np.array([2.5, 3.1])
    scf = np.array([2.0, 1.6])
    oracles = [Oracle(world, maxk=K, maxl=L, nyear=96, adrag=float(adrag[m]), scf=float(scf[m])) for m in range(2)]
    assert int(oracles[0].s("isles")) == nisl
    with Ensemble(str(tmp_path), n_members=2, perturb={"adrag": adrag, "scf": scf}) as e:
        e.set_tracer_variant("strict")
        e.run(5)
        for m, o in enumerate(oracles):
            o.run(5)
            assert np.array_equal(e.get("psi", m), o.f("psi")), (world, m)
            assert np.array_equal(e.get("ub", m), o.f("ub")), (world, m)
            assert np.array_equal(e.get("u", m), interior(o, "u")), (world, m)
        assert np.abs(e.get("psi", 0)).max() > 0.0
        nsteps = 30
        e.run(5 * (nsteps - 1))
        worst = 0.0
        for m, o in enumerate(oracles):
            o.run(5 * (nsteps - 1))
            ts = interior(o, "ts")
            scale = np.abs(ts.reshape(-1, L)).max(axis=0)
            err = np.abs(e.get("ts", m) - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
            worst = max(worst, float(err.max()))
            assert err.max() <= 1e-10 * nsteps, (world, m, float(err.max()))
        strict_psi = e.get("psi", 1)
        assert int(e.health().sum()) == 0
    print("%s (%d islands): worst per-cell relative difference of ts after %d ocean steps %.2e" % (world, nisl, nsteps, worst))
    # the production momentum path (blocked barotropic solve, forked schedule) on the same topography
    with Ensemble(str(tmp_path), n_members=2, perturb={"adrag": adrag, "scf": scf}) as e:
        e.set_tracer_variant("col")
        e.run(5 * nsteps)
        psi = e.get("psi", 1)
        assert np.abs(psi - strict_psi).max() <= 1e-9 * np.abs(strict_psi).max()
        assert int(e.health().sum()) == 0
