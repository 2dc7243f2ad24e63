"""CPU tests of the oracle itself (oracle/ is test infrastructure; parity is unpinned):
invariants the reference's own physics guarantees, known-answer checks we can derive by hand, and a
self-generated regression vector (tests/golden/, made by tools/make_golden.py) that pins the oracle
against accidental change."""
import json
import os

import numpy as np

from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I = J = 36


def grid(o, K, L):
    return (o.f("ts").reshape(K + 2, J + 2, I + 2, L), o.f("ts1").reshape(K + 2, J + 2, I + 2, L),
            o.i("k1").reshape(J + 2, I + 2))


def inventory(o, K, L):
    ts, _, k1 = grid(o, K, L)
    dz, ds = o.f("dz"), o.f("ds")
    tot = np.zeros(L)
    for k in range(1, K + 1):
        wet = (k1[1:J + 1, 1:I + 1] <= k)
        w = wet * ds[1:J + 1, None] * dz[k]
        tot += (ts[k, 1:J + 1, 1:I + 1, :] * w[..., None]).sum((0, 1))
    return tot


def test_eos_known_answers():
    o = Oracle("worbe2", maxk=8, maxl=2)
    ec = [o.s("ec%d" % q) for q in (1, 2, 3, 4)]
    rhosc = 1.0e3 * (2 * 7.2921e-5) * 0.05 * 6.37e6 / 9.81 / 5.0e3   # goldstein_lib.f90:51
    assert np.isclose(ec[0], -0.0559 / rhosc, rtol=1e-15) and np.isclose(ec[1], 0.7968 / rhosc, rtol=1e-15)
    # initial state: T=5, S=0 everywhere wet -> rho = ec1*5 + ec3*25 + ec4*125 (goldstein.f90:3056)
    rho = o.f("rho").reshape(9, J + 2, I + 2)
    k1 = o.i("k1").reshape(J + 2, I + 2)
    want = ec[0] * 5 + ec[1] * 0 + ec[2] * 25 + ec[3] * 125
    assert np.allclose(rho[8][k1 <= 8], want, rtol=1e-15)


def test_grid_known_answers():
    o = Oracle("worbe2", maxk=8, maxl=2)
    sv, s, dz, zw = o.f("sv"), o.f("s"), o.f("dz"), o.f("zw")
    assert sv[0] == -1.0 and abs(sv[J] - 1.0) < 1e-15
    assert np.allclose(np.diff(sv[:J + 1]), 2.0 / J, rtol=1e-14)                 # uniform in sin(lat)
    assert np.allclose(s[1:J + 1], sv[1:J + 1] - 1.0 / J, rtol=1e-14)
    assert abs(dz[1:9].sum() - 1.0) < 1e-14 and abs(zw[0] + 1.0) < 1e-14          # depth scaled to 1
    assert np.all(np.diff(dz[1:9]) < 0)                                          # thicker towards the bottom
    assert int(o.s("isles")) == 1 and int(o.s("ntot")) == 6210                   # 6210 wet cells


def test_uniform_passive_tracer_stays_uniform_and_inventories_are_conserved():
    K, L = 8, 4
    o = Oracle("worbe2", maxk=K, maxl=L)
    o.run(250)  # 50 ocean steps: non-trivial, non-divergent velocity field
    ts, ts1, k1 = grid(o, K, L)
    ts[..., 2] = 1.0
    ts1[..., 2] = 1.0
    kk = np.arange(K + 2)[:, None, None]
    ts[..., 3] = 2.0 + np.sin(kk) * np.cos(np.arange(J + 2))[None, :, None] + 0.1 * np.arange(I + 2)[None, None, :] % 1.3
    ts[:, :, 0, 3] = ts[:, :, I, 3]
    ts[:, :, I + 1, 3] = ts[:, :, 1, 3]
    ts1[..., 3] = ts[..., 3]
    ts[K + 1] = 0.0  # no surface flux
    ts1[K + 1] = 0.0
    before = inventory(o, K, L)
    o.call("tstepo")
    after = inventory(o, K, L)
    wet3 = (k1[None, 1:J + 1, 1:I + 1] <= np.arange(1, K + 1)[:, None, None])
    assert np.abs(ts[1:K + 1, 1:J + 1, 1:I + 1, 2][wet3] - 1.0).max() < 5e-13    # continuity: w closes the divergence
    assert np.all(np.abs(after - before) <= 1e-13 * np.maximum(np.abs(before), 1e-3))


def test_convection_removes_static_instability():
    K, L = 8, 2
    o = Oracle("worbe2", maxk=K, maxl=L)
    ts, ts1, k1 = grid(o, K, L)
    rho = o.f("rho").reshape(K + 1, J + 2, I + 2)
    # cold (dense) water on top of warm water in every column
    for k in range(1, K + 1):
        ts[k, :, :, 0] = 2.0 + 1.5 * k * (k1 <= k) * -1.0 + 20.0
    for k in range(1, K + 1):
        t = ts[k, :, :, 0]
        rho[k] = o.s("ec1") * t + o.s("ec3") * t * t + o.s("ec4") * t * t * t
    o.call("co")
    for j in range(1, J + 1):
        for i in range(1, I + 1):
            if k1[j, i] <= K:
                col = rho[k1[j, i]:K + 1, j, i]
                assert np.all(np.diff(col) <= 1e-18), (i, j, col)   # density does not increase upwards
    assert o.f("cost").sum() > 0


def test_embm_uniform_temperature_is_a_fixed_point_of_pure_diffusion():
    o = Oracle("worbe2", maxk=8, maxl=2)
    tq = o.f("tq").reshape(J, I, 2)
    tq[..., 0] = 7.5
    o.f("tq1")[:] = o.f("tq")
    o.f("tqa")[:] = 0.0
    before = tq[..., 0].copy()
    o.call("tstipa")   # betaz(1)=betam(1)=0: heat is purely diffused (embm-defaults.nml)
    assert np.abs(o.f("tq").reshape(J, I, 2)[..., 0] - before).max() < 1e-12


def test_deterministic_and_member_independent():
    a = Oracle("worbe2", maxk=8, maxl=2, diff1=1800.0)
    b = Oracle("worbe2", maxk=8, maxl=2, diff1=1800.0)
    c = Oracle("worbe2", maxk=8, maxl=2)
    for o in (a, b, c):
        o.run(100)
    assert np.array_equal(a.f("ts"), b.f("ts")) and np.array_equal(a.f("tq"), b.f("tq"))
    assert not np.array_equal(a.f("ts"), c.f("ts"))


def test_self_generated_golden_vector():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_eb_go_gs_36x36x8_1yr.json")))
    o = Oracle("worbe2", maxk=8, maxl=2, nyear=100)
    o.run(500)
    got = dict(T=float(inventory(o, 8, 2)[0]), S=float(inventory(o, 8, 2)[1]), tq_sum=float(o.f("tq").sum()),
               psi_min=float(o.f("psi").min()), psi_max=float(o.f("psi").max()), ice=float(o.f("varice").sum()),
               cost=float(o.f("cost").sum()))
    for k, v in g["values"].items():
        assert np.isclose(got[k], v, rtol=1e-12, atol=1e-300), (k, got[k], v)


def test_ediff_option():
    """iediff (SUBROUTINE ediff, goldstein.f90:2936-3044; tstepo_flux :2496-2515): with ediff0 = diff(2) the stratification-
    dependent part vanishes and the run is the constant-diffusivity one "except for rounding differences" (the reference's own
    words, :2890-2893); with ediff0 < diff(2) and ediffpow2 = 1 the diffusivity is ediff0 + ediff1(k) / N^2-like and capped by
    diffmax = dz^2 / (16 dt); the exponent cases 0, 1, 1/2 take the reference's pow-free branches."""
    base = dict(world="worbe2", maxk=8, maxl=2, nyear=100)
    o0 = Oracle(**base)
    o1 = Oracle(iediff=1, ediff0=1.0e-5, **base)
    o2 = Oracle(iediff=1, ediff0=0.3e-5, ediffpow2=1.0, **base)
    o3 = Oracle(iediff=1, ediff0=0.3e-5, ediffpow2=0.5, **base)
    o4 = Oracle(iediff=1, ediff0=0.3e-5, ediffpow2=0.5 + 1e-5, **base)     # the general pow() branch next to the sqrt branch
    for o in (o0, o1, o2, o3, o4):
        o.run(5 * 100)
    t = [o.f("ts").copy() for o in (o0, o1, o2, o3, o4)]
    assert np.abs(t[1] - t[0]).max() <= 1e-12 and not np.array_equal(t[2], t[0])
    assert 1e-7 < np.abs(t[2] - t[0]).max() < 1.0
    assert 0.0 < np.abs(t[4] - t[3]).max() < 1e-5
    assert np.isfinite(t[2]).all() and np.isfinite(t[3]).all()


def test_thermobaric_eos_option():
    """ieos = 1 (goldstein.f90:1066-1075, 3048-3082, 2692-2730, 2396-2408): rho carries ec(5) * T * z with ec(5) = 2.5e-5 dsc / rhosc,
    evaluated at the level's own depth zro(k) after every tstepo; cold water is relatively denser at depth, so the convective
    adjustment compares the two boxes at their interface and mixes a different set of columns than the depth-independent one."""
    base = dict(world="worbe2", maxk=8, maxl=2, nyear=100)
    o0, o1 = Oracle(**base), Oracle(ieos=1, **base)
    for o in (o0, o1):
        o.run(5 * 200)
    K, J, I = 8, 36, 36
    ts = o1.f("ts").reshape(K + 2, J + 2, I + 2, 2)[1:K + 1, 1:J + 1, 1:I + 1]
    rho = o1.f("rho").reshape(K + 1, J + 2, I + 2)[1:, 1:J + 1, 1:I + 1]
    k1 = o1.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
    rhosc = 1.0e3 * (2 * 7.2921e-5) * 0.05 * 6.37e6 / 9.81 / 5.0e3
    e1, e2, e3, e4, e5 = -0.0559 / rhosc, 0.7968 / rhosc, -0.0063 / rhosc, 3.7315e-5 / rhosc, 2.5e-5 * 5.0e3 / rhosc
    zro = o1.f("zro")[1:K + 1]
    t, s_ = ts[..., 0], ts[..., 1]
    want = e1 * t + e2 * s_ + e3 * (t * t) + e4 * (t * t * t) + e5 * t * zro[:, None, None]
    assert np.array_equal(rho[wet], want[wet])
    assert np.abs(o1.f("ts") - o0.f("ts")).max() > 1e-3 and np.isfinite(o1.f("ts")).all()
    assert o1.f("cost").sum() != o0.f("cost").sum()


def test_mueller_convection_option():
    """iconv = 1 (coshuffle + co, goldstein.f90:2667-2672, 2781-2841): moving the surface box down and the boxes it passes up by its
    thickness conserves every tracer's column inventory; afterwards no level is denser than the one below it (co has removed what
    the shuffle left); the convection diagnostic is a depth in metres (dsc * zw), 0 where nothing convected."""
    K, J, I = 8, 36, 36
    o = Oracle("worbe2", maxk=K, maxl=2, nyear=100, iconv=1)
    o.run(5 * 150)
    o.call("tstepo_flux")
    dz = o.f("dz")[1:K + 1]
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = (np.arange(1, K + 1)[:, None, None] >= k1[None])[..., None]
    inv = lambda: (np.where(wet, o.f("ts").reshape(K + 2, J + 2, I + 2, 2)[1:K + 1, 1:J + 1, 1:I + 1], 0.0) * dz[:, None, None, None]).sum(axis=0)
    before = inv()
    o.call("co")
    after = inv()
    assert np.abs(after - before).max() <= 1e-13 * np.abs(before).max()
    rho = o.f("rho").reshape(K + 1, J + 2, I + 2)[1:, 1:J + 1, 1:I + 1]
    both = wet[1:, ..., 0] & wet[:-1, ..., 0]
    assert np.all((rho[1:] <= rho[:-1])[both])
    cost = o.f("cost").reshape(J, I)
    assert cost.min() < -100.0 and cost.max() <= 0.0 and np.all(cost[k1 > K] == 0.0)
