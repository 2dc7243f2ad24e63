"""Oracle restatement of the Kraus-Turner mixed-layer scheme (imld = 1: tstepo goldstein.f90:2294-2390, SUBROUTINE krausturner
:3337-3442, wind energy input :112-141, tables :1013-1027, 1675-1686) -- CPU checks of the restatement itself:
the scheme only redistributes tracer within a column, so what it adds to the convective adjustment conserves every column
inventory; the diagnosed depth lies between the surface and the column's bottom; imld = 0 is untouched."""
import numpy as np

from oracle_lib import Oracle

I = J = 36
K, L = 8, 2


def _cols(o):
    ts = o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :]
    dz = o.f("dz")[1:K + 1]
    return (ts * dz[:, None, None, None]).sum(axis=0)


def test_krausturner_conserves_column_inventories_and_bounds_the_depth():
    a = Oracle("worbe2", maxk=K, maxl=L, nyear=100, imld=1)
    b = Oracle("worbe2", maxk=K, maxl=L, nyear=100, imld=0)
    a.run(5 * 60)
    # twin without the scheme, from the same state: flux + convective adjustment are common to both
    for n in ("ts", "ts1", "rho", "u", "u1", "cost"):
        b.f(n)[:] = a.f(n)
    before = a.f("ts").copy()
    a.call("tstepo")
    b.call("tstepo")
    assert not np.array_equal(a.f("ts"), b.f("ts"))                       # the scheme acted
    inv_a, inv_b = _cols(a), _cols(b)
    assert np.abs(inv_a - inv_b).max() <= 1e-13 * np.abs(inv_b).max()     # ... and moved tracer only within columns
    k1 = a.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    zw = a.f("zw")
    mld = a.f("mld").reshape(J, I)
    wet = k1 <= K
    assert (mld[wet] <= 0.0).all() and (mld[wet] >= zw[k1[wet] - 1] - 1e-12).all()
    assert (mld[~wet] == 0.0).all()
    mldk = a.i("mldk").reshape(J, I)
    assert ((mldk[wet] >= k1[wet]) & (mldk[wet] <= K)).all()
    # the mixed layer is homogeneous above the level holding its base, for both tracers
    ts = a.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :]
    emix = a.f("mldemix").reshape(J, I)
    n_checked = 0
    for j in range(J):
        for i in range(I):
            if wet[j, i] and emix[j, i] > 0 and mldk[j, i] < K - 1:
                col = ts[mldk[j, i]:K, j, i, :]                           # levels mldk+1 .. K (0-based slice)
                assert np.abs(col - col[-1]).max() == 0.0
                n_checked += 1
    assert n_checked > 20
    assert before.shape == a.f("ts").shape


def test_wind_energy_input_follows_the_stress():
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, imld=1)
    o.run(5)
    ke = o.f("mldketau").reshape(J, I)
    tau = o.f("tau").reshape(J, I, 2)
    assert (ke >= 0).all() and ke.max() > 0
    for j in (0, 1, J - 2, J - 1):                                        # polar rows carry their zonal mean
        assert np.ptp(ke[j]) == 0.0
    j, i = 17, 9
    tv4 = (tau[j, i, 0] + tau[j, i - 1, 0]) * 0.5
    tv2 = (tau[j, i, 1] + tau[j - 1, i, 1]) * 0.5
    r = np.sqrt(np.sqrt(tv4 * tv4 + tv2 * tv2))
    assert ke[j, i] == 2.5 * (r * r * r)


def test_self_generated_regression_vector():
    """One model year with imld = 1 against numbers this oracle produced when the restatement was written (tests/golden/, self-generated:
    it pins the oracle against accidental change, not against gfortran)."""
    import json
    import os
    from test_oracle import inventory
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_eb_go_gs_36x36x8_imld_1yr.json")))
    o = Oracle("worbe2", maxk=K, maxl=L, nyear=100, imld=1)
    o.run(500)
    inv = inventory(o, K, L)
    got = dict(T=float(inv[0]), S=float(inv[1]), mld_sum=float(o.f("mld").sum()), mld_min=float(o.f("mld").min()),
               mldk_sum=int(o.i("mldk").sum()), cost=float(o.f("cost").sum()), psi_max=float(o.f("psi").max()))
    for k, v in g["values"].items():
        assert np.isclose(got[k], v, rtol=1e-10, atol=1e-300), (k, got[k], v)


def test_self_generated_regression_vector_with_biogem():
    """... and one model year of the BIOGEM configuration with imld = 1: export production spread over the mixed layer
    (sub_calc_bio_uptake, biogem_box.f90:423-430, 1194-1378; cgo_biogem.c surface_column)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_eb_go_gs_ac_bg_36x36x16_imld_1yr.json")))
    o = Oracle(world="worjh2", maxk=16, maxl=16, nyear=96, imld=1)
    o.biogem_setup()
    o.run(480)
    ocn, part = o.f("ocn").reshape(16, J, I, 16), o.f("bio_part").reshape(16, J, I, 9)
    got = dict(DIC_sum=float(ocn[..., 2].sum()), PO4_surf=float(ocn[15, :, :, 5].sum()), O2_sum=float(ocn[..., 6].sum()),
               POC_lev15=float(part[14, :, :, 0].sum()), mld_sum=float(o.f("bg_mld").sum()),
               atm_pCO2=float(o.f("atm").reshape(J * I, 8)[:, 2].mean()))
    assert got["POC_lev15"] > 0                      # production below the top level: the mixed layer reaches there
    for k, v in g["values"].items():
        assert np.isclose(got[k], v, rtol=1e-9, atol=1e-300), (k, got[k], v)
