"""the Kraus-Turner mixed-layer scheme (imld = 1; tstepo goldstein.f90:2294-2390, SUBROUTINE krausturner
:3337-3442, wind energy input :112-141) on the device against the oracle -- run with -m gpu on a B200.

'strict' tracer variant: k_mld_pre / k_mld_save / k_mld_kt keep the reference's operation order (-fmad=false), so one tstepo from
the oracle's state is BIT-EXACT in ts, rho, cost and the diagnosed depth, with passive tracers, and together with ieos = 1 and
iconv = 1; a run from the initial state stays within 1e-10 per step and cell."""
import ctypes as C

import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_parity import CFG, add_passive_tracers, bits_equal, inject, interior

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,extra", [("A", {}), ("B", {}), ("A", {"ieos": 1, "iconv": 1})])
def test_mld_tracer_step_bit_exact(built, tmp_path, key, extra):
    cfg, okw = CFG[key]
    kw = dict(imld=1, **extra)
    o = Oracle(**okw, **kw)
    o.run(5 * 120)
    if okw["maxl"] > 2:
        add_passive_tracers(o)
    materialise(str(tmp_path), cfg, overrides={"go_" + k: v for k, v in kw.items()})
    with Ensemble(str(tmp_path), n_members=3) as e:
        assert bits_equal(e.get("mldketau", 1), o.f("mldketau"))          # host initialisation of the wind energy input
        for m in range(3):
            inject(e, o, m)
            e.put("mld", o.f("mld"), m)
        e.set_tracer_variant("strict")
        nsteps = 3
        for _ in range(nsteps):
            e._ck(e.L.cg_tracer_step(e.h, 1))
            o.call("tstepo")
        for m in range(3):
            ts = e.get("ts", m)
            assert bits_equal(ts, interior(o, "ts")), "ts max diff %.3e" % np.max(np.abs(ts - interior(o, "ts")))
            assert bits_equal(e.get("rho", m), interior(o, "rho"))
            assert bits_equal(e.get("cost", m), o.f("cost"))
            assert bits_equal(e.get("mld", m), o.f("mld"))
            assert bits_equal(e.get("mldpelayer1", m), np.where(o.i("k1").reshape(38, 38)[1:37, 1:37].ravel() <= okw["maxk"],
                                                              o.f("mldpelayer1"), 0.0))
        out = np.zeros(36 * 36)
        e._ck(e.L.cg_goldstein_mldta(e.h, 2, out.ctypes.data_as(C.POINTER(C.c_double))))
        assert bits_equal(out, -5000.0 * o.f("mld"))
    emix = o.f("mldemix")
    assert (emix > 0).sum() > 50 and (o.f("mld") < o.f("zw")[okw["maxk"] - 1]).sum() > 20   # krausturner deepened some columns


def test_mld_run_matches_oracle(built, tmp_path):
    cfg, okw = CFG["A"]
    K, L = okw["maxk"], okw["maxl"]
    nsteps = 30
    diff1 = np.array([2000.0, 1700.0, 2300.0])
    scf = np.array([2.0, 1.8, 2.3])                    # the wind energy input scales with scf: per-member mldketau
    materialise(str(tmp_path), cfg, overrides={"go_imld": 1})
    with Ensemble(str(tmp_path), n_members=3, perturb={"diff1": diff1, "scf": scf}) as e:
        e.set_tracer_variant("strict")
        e.run(5 * nsteps)
        got = [{n: e.get(n, m) for n in ("ts", "mld")} for m in range(3)]
        e.set_tracer_variant("col")                    # the column kernel does not take the option: the generic kernels run
        assert e.tracer_variant_active() == "fast"
        e.run(5)
        assert int(e.health().sum()) == 0
    for m in range(3):
        o = Oracle(**okw, imld=1, diff1=float(diff1[m]), scf=float(scf[m]))
        o.run(5 * nsteps)
        ts = interior(o, "ts")
        scale = np.abs(ts.reshape(-1, L)).max(axis=0)
        err = np.abs(got[m]["ts"] - ts).reshape(-1, L) / np.maximum(np.abs(ts).reshape(-1, L), 1e-3 * scale)
        dm = np.abs(got[m]["mld"] - o.f("mld"))
        print("member %d: worst per-cell relative difference after %d ocean steps %.2e, mixed-layer depth %.2e (of dsc)" %
              (m, nsteps, float(err.max()), float(dm.max())))
        assert err.max() <= 1e-10 * nsteps, (m, float(err.max()))
        assert dm.max() <= 1e-9
    o0 = Oracle(**okw, diff1=float(diff1[0]), scf=float(scf[0]))
    o0.run(5 * nsteps)
    assert np.abs(interior(o0, "ts") - got[0]["ts"]).max() > 1e-6     # the option is acting


def test_mld_with_biogem_matches_oracle(built, tmp_path):
    """imld = 1 under BIOGEM: biogem_climate takes the mixed-layer depth over (go_mldta, biogem.f90:2183) and sub_calc_bio_uptake
    spreads export production, its DOM fraction and the nutrient uptake over the levels k_mld .. n_k (biogem_box.f90:423-430,
    1186-1378) -- bg_k_mld in the sweep / packets and cells kernels.  Ten BIOGEM blocks from the initial state against the oracle,
    per cell; the mixed layer reaches below the top level in most columns from the first block on."""
    from test_gpu_biogem import CFG, OKW, compare, I, J, K, LS
    materialise(str(tmp_path), CFG, overrides={"go_imld": 1})
    scf = np.array([2.0, 1.8])
    oracles = []
    for m in range(2):
        o = Oracle(**OKW, imld=1, scf=float(scf[m]))
        o.biogem_setup()
        oracles.append(o)
    with Ensemble(str(tmp_path), n_members=2, perturb={"scf": scf}) as e:
        e.set_tracer_variant("strict")
        e.run(10)
        for o in oracles:
            o.run(10)
        compare(e, oracles, 1e-12, "imld + BIOGEM, 1 block")
        e.run(90)
        for o in oracles:
            o.run(90)
        compare(e, oracles, 1e-6, "imld + BIOGEM, 10 blocks")     # the bar of test_gpu_biogem.py for ten blocks from the neutrally stable initial state
        for m, o in enumerate(oracles):
            assert np.abs(e.get("bg_mld", m) - o.f("bg_mld")).max() <= 1e-6           # metres
        assert int(e.health().sum()) == 0
    o = oracles[0]
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    deep = (o.f("bg_mld").reshape(J, I) > 100.0) & (k1 < K)                           # deeper than the top level (80.8 m)
    part = o.f("bio_part").reshape(K, J, I, LS)
    assert deep.sum() > 200 and (part[K - 2, :, :, 0][deep] > 0).sum() > 200         # POC produced in the second level there
    o0 = Oracle(**OKW)
    o0.biogem_setup()
    o0.run(100)
    assert np.abs(o0.f("ocn") - o.f("ocn")).max() > 1e-9                             # and the run differs from imld = 0
