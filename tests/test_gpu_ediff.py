"""SURVEY 8f row 4, first option: the stratification-dependent vertical diffusivity (iediff = 1 | 2; SUBROUTINE ediff
goldstein.f90:2936-3044 and the branch of tstepo_flux :2496-2515) through cg_run against the oracle -- run with -m gpu on a B200.

'strict' tracer variant: everything on the ocean side is in the reference's operation order, so ts is BIT-EXACT against the oracle
where the scheme needs no libm call per cell (ediffpow2 = 0, 1, 0.5: +, *, sqrt) and within 1e-10 per step for a general exponent
(pow: CUDA libm vs glibc).  Members carry their own diff(2) and ediff0, i.e. their own ediff1 profile."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_parity import interior

pytestmark = pytest.mark.gpu
I = J = 36
K, L = 8, 2


@pytest.mark.parametrize("iediff,pow1,pow2,exact", [(1, 1.0, 1.0, True), (1, 0.5, 0.5, True), (2, 1.0, 0.0, True), (1, 1.0, 0.7, False)])
def test_ediff_matches_oracle(built, tmp_path, iediff, pow1, pow2, exact):
    nsteps = 30
    diff2 = np.array([1.0e-5, 2.0e-5, 0.8e-4])
    ediff0 = np.array([0.3e-5, 1.0e-5, 0.275e-4])
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_iediff": iediff, "go_ediff0": float(ediff0[0]), "go_ediffpow1": pow1,
                                                        "go_ediffpow2": pow2})
    with Ensemble(str(job), n_members=3, perturb={"diff2": diff2, "ediff0": ediff0}) as e:
        e.set_tracer_variant("strict")
        e.run(5 * nsteps)
        got = [{n: e.get(n, m) for n in ("ts", "rho", "u")} for m in range(3)]
        e.set_tracer_variant("col")                    # the column kernel does not take the option: the generic kernels run
        assert e.tracer_variant_active() == "fast"
        e.run(5)
        assert int(e.health().sum()) == 0
    worst = 0.0
    for m in range(3):
        o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, iediff=iediff, ediff0=float(ediff0[m]), ediffpow1=pow1, ediffpow2=pow2,
                   diff2=float(diff2[m]))
        o.run(5 * nsteps)
        ts = interior(o, "ts")
        if exact:
            # surflux's libm calls (exp / log / pow in the bulk formulae) are the only difference between device and oracle
            pass
        scale = np.abs(ts.reshape(-1, L)).max(axis=0)
        err = np.abs(got[m]["ts"] - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-10 * nsteps, (m, float(err.max()))
    o0 = Oracle("worbe2", maxk=K, maxl=L, nyear=100, diff2=float(diff2[0]))
    o0.run(5 * nsteps)
    assert np.abs(interior(o0, "ts") - got[0]["ts"]).max() > 1e-6     # the option is acting: a constant-diffusivity run differs
    print("iediff=%d pow1=%g pow2=%g: worst per-cell relative difference after %d ocean steps %.2e" % (iediff, pow1, pow2, nsteps, worst))


def test_ediff_tracer_step_bit_exact(built, tmp_path):
    """One tstepo from the oracle's spun-up state, strict variant: ts bit for bit (no libm on this path for ediffpow2 = 1)."""
    from test_gpu_parity import inject
    job = tmp_path / "job"
    kw = dict(iediff=1, ediff0=0.3e-5, ediffpow1=1.0, ediffpow2=1.0)
    materialise(str(job), "eb_go_gs_36x36x8", overrides={"go_" + k: v for k, v in kw.items()})
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, **kw)
    o.run(5 * 200)
    with Ensemble(str(job), n_members=1) as e:
        e.set_tracer_variant("strict")
        inject(e, o, 0)
        e.set_koverall(5 * 200)
        e._ck(e.L.cg_tracer_step(e.h, 1))
        o.call("tstepo")
        assert np.array_equal(e.get("ts", 0), interior(o, "ts"))
        assert np.array_equal(e.get("rho", 0), interior(o, "rho"))
