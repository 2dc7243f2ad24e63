"""GPU parity of the BIOGEM + ATCHEM path (run with -m gpu on a B200): the frozen eb_go_gs_ac_bg configuration
through cg_run (device) against oracle/cgo_biogem.c on identical inputs, from the initial state.

Bars (BASELINE.json north_star): <= 1e-10 relative per step on ts and the BIOGEM tracers.  The physics runs in the
'strict' variant (bit-exact), so every difference comes from the transcendental functions of the carbonate
chemistry / gas exchange (CUDA libm vs glibc, ~1 ulp)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

CFG = "eb_go_gs_ac_bg_36x36x16"
OKW = dict(world="worjh2", maxk=16, maxl=16, nyear=96)
I = J = 36
K = L = 16
LS, LA = 9, 8
PERT = {"par_bio_k0_PO4": 2.3e-6, "par_bio_remin_POC_eL1": 430.0, "par_bio_red_POC_CaCO3": 0.17, "diff1": 2200.0,
        "scf": 1.9}


FLOOR = 1e-3


def rel(a, b, floor):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def wet_masks(o):
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet3 = (np.arange(1, K + 1)[:, None, None] >= k1[None])
    return k1, wet3


@pytest.fixture(scope="module")
def pair(built, tmp_path_factory):
    d = tmp_path_factory.mktemp("bgjob")
    materialise(str(d), CFG)
    base = {k: {"par_bio_k0_PO4": 2.0e-6, "par_bio_remin_POC_eL1": 500.0, "par_bio_red_POC_CaCO3": 0.2, "diff1": 2000.0,
                "scf": 2.0}[k] for k in PERT}
    pert = {k: np.array([base[k], PERT[k]]) for k in PERT}
    oracles = []
    for m in range(2):
        kw = dict(OKW)
        kw.update({k: float(pert[k][m]) for k in ("diff1", "scf")})
        o = Oracle(**kw)
        o.biogem_setup(**{k: float(pert[k][m]) for k in PERT if k.startswith("par_bio")})
        oracles.append(o)
    e = Ensemble(str(d), n_members=2, perturb=pert)
    e.set_tracer_variant("strict")
    yield e, oracles
    e.close()


def compare(e, oracles, tol, what):
    worst = {}
    for m, o in enumerate(oracles):
        k1, wet3 = wet_masks(o)
        wl = np.repeat(wet3.ravel(), L)
        ws = np.repeat(wet3.ravel(), LS)
        ocn_o = o.f("ocn")
        ocn_d = e.get("ocn", m)
        ts_o = o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :].ravel()
        ts_d = e.get("ts", m)
        # per CELL: relative to the cell's own value, floored at 1e-3 of the tracer's largest value (temperature in degC and the
        # salinity anomaly pass through zero; particles and DOM fall by orders of magnitude with depth)
        scale = np.abs(ocn_o.reshape(-1, L)[wet3.ravel()]).max(axis=0)
        floor = np.tile(np.maximum(FLOOR * scale, 1e-300), wet3.size)
        worst["ocn"] = max(worst.get("ocn", 0), rel(ocn_d[wl], ocn_o[wl], floor[wl]))
        worst["ts"] = max(worst.get("ts", 0), rel(ts_d[wl], ts_o[wl], floor[wl]))
        po, pd = o.f("bio_part"), e.get("bio_part", m)
        pscale = np.tile(np.maximum(FLOOR * np.abs(po.reshape(-1, LS)).max(axis=0), 1e-300), wet3.size)
        worst["bio_part"] = max(worst.get("bio_part", 0), rel(pd[ws], po[ws], pscale[ws]))
        wet2 = (k1 <= K).ravel()
        Ho = o.f("carb").reshape(J * I, -1)[:, 0]
        worst["carbH"] = max(worst.get("carbH", 0), rel(e.get("carbH", m)[wet2], Ho[wet2], 1e-300))
        ao, ad = o.f("atm").reshape(J * I, LA), e.get("atm", m).reshape(J * I, LA)
        for la in (2, 3, 4, 5):
            worst["atm"] = max(worst.get("atm", 0), rel(ad[:, la], ao[:, la], 1e-300))
        so = o.f("bio_settle").reshape(K, J, I, LS)
        sd = e.get("settle_k1", m).reshape(J, I, LS)
        kk = np.clip(k1 - 1, 0, K - 1)
        so1 = np.take_along_axis(so, kk[None, :, :, None], axis=0)[0]
        sscale = np.maximum(np.abs(so1).reshape(-1, LS).max(axis=0), 1e-300)
        worst["settle"] = max(worst.get("settle", 0), float(np.max(np.abs(sd - so1)[k1 <= K] / sscale)))
    print(what, {k: "%.2e" % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= tol, (what, k, v)
    return worst


def test_biogem_initial_state(pair):
    """initialise_biogem / initialise_atchem: ocn, salinity-normalised ts, pH seed, atmosphere."""
    e, oracles = pair
    compare(e, oracles, 1e-13, "init")
    for m, o in enumerate(oracles):
        _, wet3 = wet_masks(o)
        wl = np.repeat(wet3.ravel(), L)
        # everything but the pH solve is transcendental-free: bit-exact
        assert np.array_equal(e.get("ocn", m)[wl], o.f("ocn")[wl])
        assert np.array_equal(e.get("atm", m), o.f("atm"))


def inject_all(e, o, m):
    """Restart member m of the device ensemble from the oracle's complete state (physics + BIOGEM + ATCHEM)."""
    from test_gpu_parity import inject
    inject(e, o, m)
    k1, _ = wet_masks(o)
    K1 = np.clip(k1 - 1, 0, K - 1)
    e.put("sst", np.stack([o.f("tstar_ocn"), o.f("sstar_ocn")], axis=-1).ravel(), m)   # Fortran shape (2,maxi,maxj)
    for n in ("ocn", "bio_part", "bg_M", "bg_rM", "atm", "sfcatm1", "sfxsumatm"):
        e.put(n, o.f(n), m)
    e.put("carbH", o.f("carb").reshape(J * I, -1)[:, 0].copy(), m)
    e.put("bg_seaice", o.f("bg_seaice"), m)
    so = o.f("bio_settle").reshape(K, J, I, LS)
    e.put("settle_k1", np.take_along_axis(so, K1[None, :, :, None], axis=0)[0].ravel(), m)


def test_biogem_step_parity_from_spun_state(built, tmp_path):
    """Per-step bar of the north star: restart the device from the oracle's state after 4 model months (particles
    in transit at every depth, sea ice present, convection active), advance both by 2 BIOGEM steps, compare."""
    materialise(str(tmp_path), CFG)
    o = Oracle(**OKW)
    o.biogem_setup()
    o.run(160)
    with Ensemble(str(tmp_path), n_members=1) as e:
        e.set_tracer_variant("strict")
        inject_all(e, o, 0)
        e.set_koverall(160)
        compare(e, [o], 0.0, "restart (bit-exact copy)")
        for step in (1, 2):
            e.run(10)
            o.run(10)
            compare(e, [o], 1e-10, "BIOGEM step %d after restart" % step)
        assert int(e.health().sum()) == 0
    assert np.abs(o.f("bio_part")).max() > 1e-8 and np.abs(o.f("bio_settle")).max() > 0.0


def test_biogem_model_steps(pair):
    """100 koverall iterations from the initial state = 20 ocean steps, 10 BIOGEM and ATCHEM steps, member 1 with
    perturbed biological parameters.  Over many steps ulp-level differences of the transcendental functions can flip a
    marginal convective adjustment for one step (the same effect test_gpu_parity.test_full_model_from_init documents
    for the physics), so the multi-step bar is the north star's drift criterion: global inventories to 1e-9 relative
    and the fields to 1e-6."""
    e, oracles = pair
    e.run(10)
    for o in oracles:
        o.run(10)
    compare(e, oracles, 1e-10, "after 1 BIOGEM step")
    e.run(90)
    for o in oracles:
        o.run(90)
    compare(e, oracles, 1e-6, "after 10 BIOGEM steps")
    assert int(e.health().sum()) == 0
    for m, o in enumerate(oracles):
        Mo = o.f("bg_M")
        for l in (2, 5, 6, 7):   # DIC, PO4, O2, ALK inventories
            inv_o = float((o.f("ocn").reshape(-1, L)[:, l] * Mo).sum())
            inv_d = float((e.get("ocn", m).reshape(-1, L)[:, l] * e.get("bg_M", m)).sum())
            assert abs(inv_d - inv_o) <= 1e-9 * abs(inv_o), (m, l, inv_d, inv_o)
    assert np.abs(oracles[0].f("bio_part")).max() > 1e-8


def test_biogem_multi_year_drift(built, tmp_path):
    """North-star drift criterion, shortened to what the CPU oracle finishes in seconds: after 3 model years from the
    initial state (fast tracer variant, the production path) the global means of T, S, DIC and O2 and the atmospheric
    pCO2 agree with the oracle to 1e-6 relative."""
    materialise(str(tmp_path), CFG)
    years = 3
    o = Oracle(**OKW)
    o.biogem_setup()
    o.run(480 * years)
    with Ensemble(str(tmp_path), n_members=1) as e:
        e.set_tracer_variant("fast")
        e.run(480 * years)
        assert int(e.health().sum()) == 0
        ocn_d = e.get("ocn", 0).reshape(-1, L)
        M_d = e.get("bg_M", 0)
        atm_d = e.get("atm", 0).reshape(-1, LA)
    ocn_o = o.f("ocn").reshape(-1, L)
    M_o = o.f("bg_M")
    for l, name in ((0, "T"), (1, "S"), (2, "DIC"), (6, "O2")):
        mo = float((ocn_o[:, l] * M_o).sum() / M_o.sum())
        md = float((ocn_d[:, l] * M_d).sum() / M_d.sum())
        print("global mean %s: oracle %.12e device %.12e rel %.2e" % (name, mo, md, abs(md - mo) / abs(mo)))
        assert abs(md - mo) <= 1e-6 * abs(mo), (name, mo, md)
    assert abs(atm_d[0, 2] - o.f("atm").reshape(-1, LA)[0, 2]) <= 1e-6 * 278e-6
