"""GPU parity of the fused tracer column kernel (variant 'col': tstepo_flux + co + SST export in one pass,
cgenie_b200/csrc/k_tracer_col.cuh) -- run with -m gpu on a B200.

Bars (BASELINE.json north_star): <= 1e-10 relative per step on ts and the BIOGEM tracers against the oracle on
identical inputs; global means within 1e-6 after a multi-year run.  The same per-thread body is checked on the host
against the oracle in tests/test_col_body_host.py."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
from test_gpu_biogem import CFG, OKW, I, J, K, L, LA, compare, inject_all

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("members,check", [(1, (0,)), (33, (0, 32)), (100, (0, 99))])
def test_col_step_parity_from_spun_state(built, tmp_path, members, check):
    """Restart chosen members (first / last warp of the 32-, 64- and 128-lane instances) from the oracle's state after
    2 model years (convection active, particles in transit; not the first months, where the uniform initial state is
    neutrally stable and mixing decisions flip on the last bit) and advance 2 BIOGEM steps = 4 ocean steps."""
    materialise(str(tmp_path), CFG)
    o = Oracle(**OKW)
    o.biogem_setup()
    o.run(960)
    with Ensemble(str(tmp_path), n_members=members) as e:
        e.set_tracer_variant("col")
        assert e.tracer_variant_active() == "col"
        for m in check:
            inject_all(e, o, m)
        e.set_koverall(960)
        for step in (1, 2):
            e.run(10)
            o.run(10)
            for m in check:
                compare(_Member(e, m), [o], 1e-10, "col, M=%d, member %d, BIOGEM step %d" % (members, m, step))
        assert int(e.health()[list(check)].sum()) == 0


class _Member:
    """view of one ensemble member as member 0 (compare() walks oracles by index)"""

    def __init__(self, e, m):
        self.e, self.m = e, m

    def get(self, name, _m):
        return self.e.get(name, self.m)


def test_col_launch_count(built, tmp_path):
    materialise(str(tmp_path), CFG)
    counts = {}
    for variant in ("fast", "col"):
        with Ensemble(str(tmp_path), n_members=2) as e:
            e.set_tracer_variant(variant)
            e.set_graphs(False)
            n0 = e.launch_count()
            e.run(5)
            counts[variant] = e.launch_count() - n0
    assert counts["fast"] - counts["col"] == 1       # flux + co + sst -> column kernel + passive-tracer mixing


def test_col_multi_year_drift(built, tmp_path):
    """North-star drift criterion (shortened to what the CPU oracle finishes in seconds): 3 model years from the
    initial state, global means of T, S, DIC, O2 and atmospheric pCO2 within 1e-6 relative of the oracle."""
    materialise(str(tmp_path), CFG)
    years = 3
    o = Oracle(**OKW)
    o.biogem_setup()
    o.run(480 * years)
    with Ensemble(str(tmp_path), n_members=1) as e:
        e.set_tracer_variant("col")
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


def test_baro_solves_agree(built, tmp_path):
    """Barotropic solve of variant 'col'.  (i) The register-resident pivot-by-pivot solve (k_baro_reg, CG_BARO_BLK=0)
    performs the same operations in the same order as the shared-memory one (k_baro_fast, variant 'fast'): after the
    momentum part of the first ocean step the stream function, the barotropic and the 3-D velocities of perturbed members
    are bit-identical.  (ii) The blocked solve (k_baro_blk, the default: 32 unknowns at a time through the precomputed
    inverse of the block's triangle) solves the same equations in another summation order: equal to rounding.
    (The tracer step runs after the momentum step, so from step 2 on the variants' own rounding enters through rho.)"""
    import os
    materialise(str(tmp_path), CFG)
    M = 5
    pert = {"adrag": np.linspace(2.0, 3.0, M), "scf": np.linspace(1.5, 2.5, M), "diff1": np.linspace(1500.0, 2500.0, M)}
    out = {}
    for variant, blk in (("fast", None), ("col", "0"), ("col", "1")):
        os.environ.pop("CG_BARO_BLK", None)
        if blk is not None:
            os.environ["CG_BARO_BLK"] = blk
        try:
            with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
                e.set_tracer_variant(variant)
                e.run(5)
                out[(variant, blk)] = {n: np.stack([e.get(n, m) for m in range(M)]) for n in ("gb", "psi", "ub", "u")}
                assert int(e.health().sum()) == 0
        finally:
            os.environ.pop("CG_BARO_BLK", None)
    for n in ("gb", "psi", "ub", "u"):
        assert np.array_equal(out[("fast", None)][n], out[("col", "0")][n]), n
    for n in ("psi", "ub", "u"):
        ref, got = out[("col", "0")][n], out[("col", "1")][n]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 1e-13, (n, err)
    assert np.abs(out[("col", "1")]["psi"]).max() > 0.0


def test_baro_four_warp_form_bit_identical(built, tmp_path):
    """k_baro_blk4 (four warps per member, warp w = accumulator w of the one-warp form's two sums, partial sums combined in the
    same order) against k_baro_blk: gb, psi, ub, u bit for bit after ocean steps of perturbed members."""
    import os
    materialise(str(tmp_path), CFG)
    M = 40
    pert = {"adrag": np.repeat(np.linspace(2.0, 3.0, 5), 8), "scf": np.linspace(1.5, 2.5, M), "diff1": np.linspace(1500.0, 2500.0, M)}
    out = {}
    for w4 in ("1", "0"):
        os.environ["CG_BARO_W4"] = w4
        try:
            with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
                e.set_tracer_variant("col")
                e.run(5 * 6)
                out[w4] = {n: np.stack([e.get(n, m) for m in (0, 17, M - 1)]) for n in ("gb", "psi", "ub", "u", "ts")}
                assert int(e.health().sum()) == 0
        finally:
            os.environ.pop("CG_BARO_W4", None)
    for n in out["1"]:
        assert np.array_equal(out["1"][n], out["0"][n]), n
    assert np.abs(out["1"]["psi"]).max() > 0.0


def test_biogem_fused_coupling_bit_identical(built, tmp_path):
    """cg_run applies biogem_tracercoupling's per-cell update inside the step_biogem kernel (the global sums are taken
    first: they do not depend on the step's anomaly).  Same expressions in the same order: every field is bit-identical
    to the run with the two separate kernels, members with perturbed biology included."""
    materialise(str(tmp_path), CFG)
    M = 3
    pert = {"par_bio_k0_PO4": np.array([2.0e-6, 1.7e-6, 2.4e-6]), "par_bio_remin_POC_eL1": np.array([500.0, 430.0, 560.0]),
            "diff1": np.array([2000.0, 1800.0, 2300.0])}
    out = {}
    for fused in (False, True):
        with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
            e.set_tracer_variant("strict")
            e.set_biogem_fusion(fused)
            e.run(60)          # 12 ocean steps, 6 BIOGEM + ATCHEM steps
            out[fused] = {n: np.stack([e.get(n, m) for m in range(M)])
                          for n in ("ts", "ocn", "bio_part", "bg_M", "bg_rM", "atm", "carbH", "settle_k1", "sfcocn1")}
            assert int(e.health().sum()) == 0
    for n in out[True]:
        assert np.array_equal(out[False][n], out[True][n]), n
    assert np.abs(out[True]["bio_part"]).max() > 0.0


def test_concurrent_schedule_bit_identical(built, tmp_path):
    """cg_run overlaps independent parts of the coupling loop on several streams (momentum next to surflux/EMBM/sea ice
    inside the captured cycle; the BIOGEM/ATCHEM block next to the head of the following cycle).  Every kernel is
    deterministic, so any missed dependency would show as a difference: the complete state after 3 model months of a
    perturbed 40-member ensemble is bit-identical to the fully serial schedule."""
    import os
    materialise(str(tmp_path), CFG)
    M = 40
    rng = np.random.default_rng(7)
    pert = {"diff1": rng.uniform(1500.0, 2500.0, M), "adrag": rng.uniform(2.0, 3.0, M), "scf": rng.uniform(1.5, 2.5, M),
            "par_bio_k0_PO4": rng.uniform(1.7e-6, 2.4e-6, M)}
    names = ("ts", "rho", "u", "psi", "tq", "varice", "ocn", "bio_part", "atm", "cost", "bg_seaice", "sst", "carbH")
    out = {}
    knobs = ("CG_NOFORK", "CG_BG_SERIAL", "CG_NOEAGER", "CG_BG_PIPE", "CG_BG_SPLIT", "CG_BG_SWEEP_EARLY", "CG_TC_AHEAD")
    # "default" = what cg_run does out of the box; "async" = the block next to the following cycle's head, step kernel at
    # its nominal place; "pipelined" = whole step kernel one block ahead; "split" = only its surface part one block ahead
    env = {"serial": {"CG_NOFORK": "1", "CG_BG_SERIAL": "1", "CG_NOEAGER": "1"}, "default": {}, "async": {"CG_BG_SPLIT": "0"},
           "pipelined": {"CG_BG_PIPE": "1", "CG_BG_SPLIT": "0"}, "split": {"CG_BG_SPLIT": "1", "CG_BG_SWEEP_EARLY": "0"},
           "split, sweep one cycle ahead": {"CG_BG_SWEEP_EARLY": "1"},
           # default takes the coupling sums over BIOGEM's own state one block ahead; this mode takes all sums in place
           "split, all coupling sums in place": {"CG_TC_AHEAD": "0"}}
    for mode in env:
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update(env[mode])
        try:
            with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
                e.set_tracer_variant("col")
                e.run(120)
                e.run(120)       # a second call: the block left pending by the first one is joined correctly
                e.run(35)        # ... and calls that end between two BIOGEM blocks
                e.run(85)
                out[mode] = {n: e.get_all(n).copy() for n in names}
                assert int(e.health().sum()) == 0
        finally:
            for k in knobs:
                os.environ.pop(k, None)
    for mode in env:
        for n in names:
            assert np.array_equal(out["serial"][n], out[mode][n]), (mode, n)


def test_module_by_module_matches_run(built, tmp_path):
    """The per-module entry points (what the Fortran shims call once per koverall iteration: a cycle of calls without host
    arrays is issued as cg_run issues it, two graph replays, when step_goldstein completes it; otherwise momentum is started
    at surflux time; BIOGEM calls run on their own stream) give bit-identical state to cg_run's schedule."""
    materialise(str(tmp_path), CFG)
    M = 6
    rng = np.random.default_rng(11)
    pert = {"diff1": rng.uniform(1500.0, 2500.0, M), "adrag": rng.uniform(2.0, 3.0, M),
            "par_bio_k0_PO4": rng.uniform(1.7e-6, 2.4e-6, M)}
    names = ("ts", "rho", "u", "psi", "tq", "varice", "ocn", "bio_part", "atm", "cost", "bg_seaice", "sst")
    out = {}
    for mode in ("run", "modules"):
        with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
            e.set_tracer_variant("col")
            e.run(100)
            if mode == "run":
                e.run(100)
            else:
                genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
                tick = int(round(1000.0 * genie_timestep))
                dts = float(2 * 5) * genie_timestep
                for k in range(101, 201):
                    if k == 151:      # a host write between two BIOGEM blocks: the surface part issued ahead is dropped
                        e.put_all("ocn", e.get_all("ocn").copy())
                    if k % 5 == 1:
                        e.surflux()
                    e.step_embm()
                    if k == 153:      # a host read in the middle of a cycle: the calls noted so far are replayed one by one
                        e.get("tq", 0)
                    if k == 163:      # a host write to an input of the momentum step after it was started early (the read replays
                        e.put("rho", e.get("rho", 2), 2)   # the noted calls): u1 is rolled back, step_goldstein repeats the step
                    if k % 5 == 0:
                        e.step_seaice()
                        if k == 175:  # ... and between the sea-ice and the ocean step
                            e.get("varice", 0)
                        if k == 185:  # go_ts / go_cost INOUT through step_goldstein (goldstein.f90:36, 99, 176): uploaded before the
                            import ctypes as C            # step, downloaded after; the early momentum step stays valid
                            from cgenie_b200._lib import D, GoldsteinIO
                            ts_io, cost_io = e.get("ts", 0).copy(), e.get("cost", 0).copy()     # the Fortran host drives member 0
                            io = GoldsteinIO()
                            io.go_ts, io.go_cost = ts_io.ctypes.data_as(D), cost_io.ctypes.data_as(D)
                            e.step_goldstein(C.byref(io))
                            assert np.array_equal(ts_io, e.get("ts", 0)) and np.array_equal(cost_io, e.get("cost", 0))
                        else:
                            e.step_goldstein()
                    if k % 10 == 0:
                        e.biogem_forcing(k * tick)
                        e.biogem_step(dts, k * tick)
                        e.biogem_tracercoupling()
                        e.biogem_climate()
                        e.atchem_step(dts)
            out[mode] = {n: e.get_all(n).copy() for n in names}
            assert int(e.health().sum()) == 0
    for n in names:
        assert np.array_equal(out["run"][n], out["modules"][n]), n


def test_ensemble_groups_match_single_handle(built, tmp_path):
    """A shard run as independent groups (EnsembleGroups: one library handle per group of <= 128 members, driven from host
    threads) gives every member the state it has in a single handle: members never interact, so the grouping is invisible."""
    from cgenie_b200 import EnsembleGroups
    materialise(str(tmp_path), CFG)
    M = 64
    rng = np.random.default_rng(5)
    pert = {"diff1": rng.uniform(1500.0, 2500.0, M), "adrag": np.repeat(rng.uniform(2.0, 3.0, 2), 32),
            "par_bio_k0_PO4": rng.uniform(1.7e-6, 2.4e-6, M)}
    names = ("ts", "rho", "u", "tq", "varice", "ocn", "bio_part", "atm")
    check = (0, 31, 32, 63)
    with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e:
        e.set_tracer_variant("col")
        e.run(200)
        one = {(n, m): e.get(n, m).copy() for n in names for m in check}
        means_one = e.global_means()
    with EnsembleGroups(str(tmp_path), n_members=M, perturb=pert, group=32) as g:
        assert len(g.parts) == 2
        g.set_tracer_variant("col")
        assert g.tracer_variant_active() == "col"
        g.run(200)
        g.synchronize()
        assert int(g.health().sum()) == 0 and g.health().shape == (M,)
        assert g.launch_count() > 0
        for n in names:
            for m in check:
                a, b = one[(n, m)], g.get(n, m)
                assert np.allclose(a, b, rtol=1e-12, atol=1e-12 * max(np.abs(a).max(), 1e-300)), (n, m, np.abs(a - b).max())
        assert np.allclose(means_one, g.global_means(), rtol=1e-12)


def test_col_century_drift(built, tmp_path):
    """North-star drift criterion at its full length: 100 model years from the initial state, the unperturbed control and
    one perturbed member (physics + biology), global means of T, S, DIC, O2 and atmospheric pCO2 within 1e-6 relative of
    the oracle (measured on B200: <= 2.1e-10).  The device needs ~4 s, the two oracle threads ~80 s."""
    import threading
    from cgenie_b200.sharding import perturbation_table
    materialise(str(tmp_path), CFG)
    years = 100
    tab = perturbation_table(4, biogem=True)
    members = (0, 3)
    ref, keep = {}, {}

    def oracle_run(m):
        kw = {k: float(v[m]) for k, v in tab.items()}
        o = Oracle(**dict(OKW, **{k: v for k, v in kw.items() if not k.startswith("par_bio")}))
        o.biogem_setup(**{k: v for k, v in kw.items() if k.startswith("par_bio")})
        o.run(480 * years)
        ref[m] = (o.f("ocn").reshape(-1, L).copy(), o.f("bg_M").copy(), o.f("atm").reshape(-1, LA).copy())
        keep[m] = o
    th = [threading.Thread(target=oracle_run, args=(m,)) for m in members]
    for t in th:
        t.start()
    with Ensemble(str(tmp_path), n_members=4, perturb=tab) as e:
        e.set_tracer_variant("col")
        e.run(480 * years)
        assert int(e.health().sum()) == 0
        dev = {m: (e.get("ocn", m).reshape(-1, L), e.get("bg_M", m), e.get("atm", m).reshape(-1, LA)) for m in members}
    for t in th:
        t.join()
    for m in members:
        assert m in ref, "oracle thread failed"
        ocn_o, M_o, atm_o = ref[m]
        ocn_d, M_d, atm_d = dev[m]
        for l, name in ((0, "T"), (1, "S"), (2, "DIC"), (6, "O2")):
            mo = float((ocn_o[:, l] * M_o).sum() / M_o.sum())
            md = float((ocn_d[:, l] * M_d).sum() / M_d.sum())
            print("member %d global mean %s: oracle %.12e device %.12e rel %.2e" % (m, name, mo, md, abs(md - mo) / abs(mo)))
            assert abs(md - mo) <= 1e-6 * abs(mo), (m, name, mo, md)
        assert abs(atm_d[0, 2] - atm_o[0, 2]) <= 1e-6 * atm_o[0, 2]
    # the per-step bar AT the state the bench times (100 model years old): both members restarted from their oracle's century
    # state inside a 128-lane ensemble, then 2 BIOGEM steps = 4 ocean steps of the production kernels against the oracle
    tab128 = perturbation_table(128, biogem=True)
    with Ensemble(str(tmp_path), n_members=128, perturb=tab128) as e:
        e.set_tracer_variant("col")
        for m in members:
            inject_all(e, keep[m], m)
        e.set_koverall(480 * years)
        for step in (1, 2):
            e.run(10)
            for m in members:
                keep[m].run(10)
                compare(_Member(e, m), [keep[m]], 1e-10, "col, century state, member %d, BIOGEM step %d" % (m, step))
        assert int(e.health()[list(members)].sum()) == 0
    for o in keep.values():
        o.close()
