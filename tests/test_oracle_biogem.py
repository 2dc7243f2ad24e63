"""CPU tests of the BIOGEM/ATCHEM oracle (oracle/cgo_biogem.c; test infrastructure, parity unpinned): element mass
balances that the reference's own audit checks (biogem.f90:1766-1785), hand-derivable known answers, and a
self-generated regression vector (tests/golden/, tools/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I = J = 36
K = L = 16
LS, LA = 9, 8
# compact indices (0-based) of the frozen configuration
DIC, DIC13, DIC14, PO4, O2, ALK, DOMC, DOMC13, DOMC14, DOMP, CA = 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12
POC, POC13, POC14, POP, CACO3 = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def bg():
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    o.biogem_setup()
    return o


def budgets(o):
    """Ocean inventories (mol) including particulates in the water column and the settled material that the closed
    system returns to the bottom cell at the next step (biogem.f90:887-925)."""
    ocn = o.f("ocn").reshape(K, J, I, L)
    part = o.f("bio_part").reshape(K, J, I, LS)
    M = o.f("bg_M").reshape(K, J, I)
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    so = o.f("bio_settle").reshape(K, J, I, LS)
    s1 = np.take_along_axis(so, np.clip(k1 - 1, 0, K - 1)[None, :, :, None], axis=0)[0] * (k1 <= K)[..., None]
    inv = lambda a: float((a * M).sum())
    P = inv(ocn[..., PO4]) + inv(ocn[..., DOMP]) + inv(part[..., POP]) + float(s1[..., POP].sum())
    Ca = inv(ocn[..., CA]) + inv(part[..., CACO3]) + float(s1[..., CACO3].sum())
    A = (inv(ocn[..., ALK]) - 16.0 * (inv(ocn[..., DOMP]) + inv(part[..., POP]) + float(s1[..., POP].sum()))
         + 2.0 * (inv(part[..., CACO3]) + float(s1[..., CACO3].sum())))
    return dict(P=P, Ca=Ca, ALKstar=A)


def test_initial_state(bg):
    o = bg
    ocn = o.f("ocn").reshape(K, J, I, L)
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
    assert np.all(ocn[..., DIC][wet] == 2.244e-3) and np.all(ocn[..., PO4][wet] == 2.159e-6)
    # isotopes: abundance = R/(1+R)*total with R = standard*(1+delta/1000) (gem_util.f90:604-617)
    R = 0.011202 * (1.0 + 0.4 / 1000.0)
    assert np.allclose(ocn[..., DIC13][wet], R / (1 + R) * 2.244e-3, rtol=1e-15)
    # salinity-normalised copy on ts: uniform S -> ts == ocn to rounding
    ts = o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1]
    assert np.allclose(ts[..., DIC][wet], 2.244e-3, rtol=1e-14)
    # initial surface carbonate system for DIC 2244 / ALK 2363 umol/kg at 5 C, S 34.9: seeded at pH 7.8, solves to
    # pH(SWS) 7.94, [CO3] ~ ALK-DIC-borate ~ 95 umol/kg, calcite saturation ~2.3, Revelle factor ~16 (cold water)
    carb = o.f("carb").reshape(J, I, -1)[k1 <= K]
    assert np.all(np.abs(-np.log10(carb[:, 0]) - 7.944) < 0.01)
    assert np.all(np.abs(carb[:, 2] - 95e-6) < 5e-6) and np.all(np.abs(carb[:, 5] - 2.26) < 0.05)
    assert np.all(np.abs(carb[:, 9] - 16.0) < 0.5)
    assert np.allclose(carb[:, 1] + carb[:, 2] + carb[:, 3], 2.244e-3, rtol=1e-12)   # CO2 + CO3 + HCO3 = DIC
    atm = o.f("atm").reshape(J, I, LA)
    assert np.all(atm[..., 2] == 278.0e-6) and np.all(atm[..., 0] == 273.15)


def test_mass_balances_over_a_year(bg):
    """P, Ca and the alkalinity combination ALK - 16 (POP + DOP) + 2 CaCO3 are conserved by uptake, DOM cycling,
    sinking, remineralisation, the closed-system sediment return and the salinity-normalised tracer coupling."""
    o = bg
    b0 = budgets(o)
    o.run(480)
    b1 = budgets(o)
    for k in b0:
        assert abs(b1[k] - b0[k]) <= 2e-11 * abs(b0[k]), (k, b0[k], b1[k])
    part = o.f("bio_part").reshape(K, J, I, LS)
    assert part[..., POC].max() > 1e-8                      # export production happened
    assert np.abs(o.f("ocn").reshape(K, J, I, L)[..., DOMC]).max() > 1e-6   # DOM pool built up
    atm = o.f("atm").reshape(J, I, LA)
    assert abs(atm[0, 0, 2] - 278e-6) < 5e-6                # pCO2 restored towards 278 ppm (tau = 0.1 yr)
    assert np.ptp(atm[..., 2]) == 0.0                       # homogenised atmosphere (atchem.f90:140-150)
    assert o.s("bg_go") == 1.0 and o.s("bg_clock_ms") == 480 * round(1000.0 * 3600.0 * 24.0 * 365.25 / 5.0 / 96)


def test_particle_flux_profile(bg):
    """After a year the POC reaching depth follows the two-fraction e-folding profile: below ~1 km the labile
    fraction (eL1 = 500 m) is gone and the recalcitrant share frac2 approaches 1."""
    o = bg
    part = o.f("bio_part").reshape(K, J, I, LS)
    f2 = part[..., 7]
    poc = part[..., POC]
    deep = (poc > 1e-12) & (np.arange(K)[:, None, None] < 6)
    assert deep.any() and f2[deep].min() > 0.3
    assert np.all((f2 >= 0) & (f2 <= 1.0 + 1e-12))


def test_regression_vector(bg):
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_eb_go_gs_ac_bg_36x36x16_1yr.json")))["values"]
    o = bg                                                   # state after exactly one model year (test above)
    ocn = o.f("ocn").reshape(K, J, I, L)
    M = o.f("bg_M").reshape(K, J, I)
    got = dict(DIC=float((ocn[..., DIC] * M).sum()), O2=float((ocn[..., O2] * M).sum()),
               PO4=float((ocn[..., PO4] * M).sum()), pCO2=float(o.f("atm").reshape(J, I, LA)[0, 0, 2]),
               POC_max=float(o.f("bio_part").reshape(K, J, I, LS)[..., POC].max()))
    for k, v in want.items():
        assert np.isclose(got[k], v, rtol=1e-9, atol=0.0), (k, got[k], v)


def test_sedgem_coupler_sums():
    """cpl_flux_ocnsed / cpl_comp_ocnsed / reinit_flux_rokocn (sedgem.f90:1029-1068, :894-937, rokgem.f90:472-480; called
    after every BIOGEM step whether or not SEDGEM runs, genie.f90:413-427): time integral of the ocean->sediment flux and the
    running mean of the bottom-water composition over the BIOGEM steps of one SEDGEM step (conv_kocn_ksedgem = 8 here)."""
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    o.biogem_setup()
    dts = float(2 * 5) * (3600.0 * 24.0 * 365.25 / 5.0 / 96)
    flux, comp, want_w = np.zeros(J * I * LS), np.zeros(J * I * L), [0, 1, 2, 3, 0, 1]
    seen = []
    for blk in range(1, 7):
        o.run(10)
        ocnstep = 10 * blk // 5
        o.L.cgo_cpl_flux_ocnsed(o.h, dts)
        o.L.cgo_cpl_comp_ocnsed(o.h, ocnstep, 2, 8)
        w = want_w[blk - 1]
        flux = flux + dts * o.f("sfxsed1")
        comp = (float(w) * comp + o.f("sfcocn1")) / float(w + 1)
        assert np.array_equal(o.f("sfxsumsed"), flux) and np.array_equal(o.f("sfcsumocn"), comp), blk
        seen.append(o.f("sfcocn1").copy())
    # the mean restarts with each SEDGEM step: after 6 BIOGEM steps it is the mean of the last two
    assert np.allclose(comp, 0.5 * (seen[4] + seen[5]), rtol=1e-15, atol=0)
    # what left the ocean for the sediments is what the closed system hands back: CaCO3 rain integrated over the run is positive
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    rain = o.f("sfxsumsed").reshape(J, I, LS)
    assert rain[..., POC][k1 <= K].min() >= 0.0 and rain[..., CACO3][k1 <= K].max() > 0.0 and np.all(rain[k1 > K] == 0.0)
    o.f("sfxsumrok1")[:] = 3.0
    o.L.cgo_reinit_flux_rokocn(o.h)
    assert np.all(o.f("sfxsumrok1") == 0.0)


def test_timeslice_diagnostics():
    """diag_biogem_timeslice (biogem.f90:2421-2699) restated: the window integrals are dtyr-weighted sums of the fields at the
    call point, the 3-D carbonate re-solve lands on the published deep-ocean ranges (pH 7.5 - 8.3, calcite under-saturated at
    depth, over-saturated at the surface), and the only thing the diagnostic changes in the model is the surface [H+] seed of the
    next step's solve (same fixed point: the trajectory moves by no more than the solver's 0.1 % tolerance allows)."""
    N_IC, N_CC = 10, 17
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    o.biogem_setup()
    ref = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    ref.biogem_setup()
    dtyr = float(2 * 5) / 5.0 / 96
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet3 = (np.arange(1, K + 1)[:, None, None] >= k1[None]).ravel()
    o.run(480)
    ref.run(480)
    acc, accp = np.zeros(K * J * I * L), np.zeros(K * J * I * LS)
    for blk in range(3):
        o.run(10)
        ref.run(10)
        o.L.cgo_biogem_slice_update(o.h)
        acc = acc + dtyr * o.f("ocn")
        accp = accp + dtyr * o.f("bio_part")
    w = np.repeat(wet3, L)
    assert np.array_equal(o.f("sl_ocn")[w], acc[w]) and np.all(o.f("sl_ocn")[~w] == 0.0)
    assert np.array_equal(o.f("sl_part")[np.repeat(wet3, LS)], accp[np.repeat(wet3, LS)])
    assert abs(o.f("sl_t")[0] - 3 * dtyr) < 1e-15
    carb = (o.f("sl_carb") / o.f("sl_t")[0]).reshape(K, J * I, N_IC)
    w3 = wet3.reshape(K, J * I)
    pH = -np.log10(carb[..., 0][w3])
    assert 7.4 < pH.min() and pH.max() < 8.4, (pH.min(), pH.max())
    assert carb[K - 1, :, 5][w3[K - 1]].min() > 1.0          # calcite saturation at the surface
    deep = carb[0, :, 5][w3[0]]                              # the deepest level: 5000 m
    assert deep.size > 0 and deep.max() < 1.2 and deep.min() > 0.3
    # CO2 + HCO3 + CO3 = DIC in every cell
    dic = (o.f("sl_ocn") / o.f("sl_t")[0]).reshape(K, J * I, L)[..., DIC]
    s = carb[..., 1] + carb[..., 2] + carb[..., 3]
    assert np.max(np.abs(s[w3] / dic[w3] - 1.0)) < 1e-12
    cc = (o.f("sl_carbconst") / o.f("sl_t")[0]).reshape(K, J * I, N_CC)
    pK1 = -np.log10(cc[K - 1, :, 0][w3[K - 1]])
    assert 5.7 < pK1.min() and pK1.max() < 6.2                # Mehrbach refit: pK1 = 5.84 at 25 degC, 6.1 at 0 degC
    # the model went on unchanged but for the surface seed
    rel = np.abs(o.f("ocn")[w] - ref.f("ocn")[w]) / np.maximum(np.abs(ref.f("ocn")[w]), 1e-3 * np.abs(ref.f("ocn")[w]).max())
    assert rel.max() < 1e-6


def test_extended_timeseries_integrals_are_consistent():
    """The oracle's restatement of the flux / export / misc integrals of diag_biogem_timeseries (cgo_biogem_sig_update, "bg_sig2";
    biogem.f90:2870-2883, 2926-2964, 3058-3062): six BIOGEM steps of the third model year -- the integrals must tie in with state the
    model keeps elsewhere (particulate export against the POC the surface layer produced, isotope ratios of the fluxes against
    those of their bulk tracers, sea ice against the cover, the overturning extrema around zero)."""
    import numpy as np
    from oracle_lib import Oracle
    o = Oracle(world="worjh2", maxk=16, maxl=16, nyear=96)
    o.biogem_setup()
    o.run(1400)                                       # the first sea ice forms at the end of the third model year
    o.L.cgo_biogem_sig_auto(o.h, 1, 1000.0)
    o.run(60)
    S, X = o.f("bg_sig"), o.f("bg_sig2")
    LS, LA = 9, 8
    assert abs(S[0] - 6.0 / 48.0) < 1e-12 and X.size == 8 + LS + 2 * LA
    fe, oa, sa = X[8:8 + LS], X[8 + LS:8 + LS + LA], X[8 + LS + LA:]
    assert X[0] > 1e9 and X[2] > 0 and 0 < X[1] / S[0] < 10.0             # m2 yr, m3 yr, m yr
    assert X[3] < 0 < X[4] and X[5] <= 0 <= X[6] and abs(X[3]) < 1 and abs(X[4]) < 1
    assert -40.0 < X[7] / S[0] < 40.0                                            # degrees C over land
    assert fe[0] > 0 and fe[4] > 0 and fe[3] > 0
    assert abs(fe[0] / fe[3] - 106.0) < 1e-6 * 106.0                      # POC : POP leave the surface in the Redfield ratio
    assert 0.0105 < fe[1] / fe[0] < 0.0112 and 0.0108 < fe[5] / fe[4] < 0.0113   # 13C fractions of POC and CaCO3
    assert oa[0] == 0 and oa[1] == 0 and sa[0] == 0 and sa[1] == 0         # temperature and humidity carry no flux
    assert sa[2] != 0 and sa[5] != 0 and 0.010 < sa[3] / sa[2] < 0.012     # CO2, O2 exchange; 13C fraction of the CO2 flux
    # focnatm is the restoring forcing net of the gas exchange: with pCO2 restored it differs from the pure exchange
    assert oa[2] != sa[2]
    # ... and O2 is not restored: the interface flux (conv_yr_s * A * sfxatm1) IS the gas exchange (mol yr-1), two routes to one number
    assert abs(oa[5] - sa[5]) <= 1e-12 * abs(sa[5])
