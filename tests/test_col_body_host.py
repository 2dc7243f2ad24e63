"""Fused tracer column kernel (cgenie_b200/csrc/k_tracer_col.cuh): its per-thread body compiled for the HOST and checked
against the oracle's tstepo (tstepo_flux + co, goldstein.f90:2280-2777) on spun-up eb_go_gs_ac_bg states.  No GPU needed;
the GPU parity of the same code is tests/test_gpu_col.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
I = J = 36
K = L = 16
MS = 32


def _lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "col_host")])
    lib = C.CDLL(os.path.join(HERE, "col_host", "libcol_host.so"))
    lib.col_host_step.restype = C.c_int
    return lib


def _state(o):
    """device-layout (member axis of length 1) views of one oracle's tracer-step inputs"""
    ts1 = o.f("ts1").reshape(K + 2, J + 2, I + 2, L)
    u = o.f("u").reshape(K, J + 1, I + 1, 3)
    rho = o.f("rho").reshape(K + 1, J + 2, I + 2)
    return dict(ts=ts1[1:K + 1, 1:J + 1, 1:I + 1, :].copy(), u=u[:, 1:, 1:, :].copy(),
                tsflux=np.ascontiguousarray(ts1[K + 1, 1:J + 1, 1:I + 1, 0:2].transpose(2, 0, 1)),
                rho=rho[1:, 1:J + 1, 1:I + 1].copy(), cost=o.f("cost").reshape(J, I).copy())


def _after(o):
    ts = o.f("ts").reshape(K + 2, J + 2, I + 2, L)
    rho = o.f("rho").reshape(K + 1, J + 2, I + 2)
    return ts[1:K + 1, 1:J + 1, 1:I + 1, :].copy(), rho[1:, 1:J + 1, 1:I + 1].copy(), o.f("cost").reshape(J, I).copy()


# 5: warp-tile column kernel (one warp per block, lane-parallel row copies) + co;
# 4: round-1 flux kernel + stability flag, decisions-only convection kernel + one thread per passive tracer;
# 3: pipelined column kernel (coefficients one level ahead; production) + stability flag + co on flagged member-columns;
# 2: split column kernel (two threads per member-column) + co; 1: T,S pre-pass + mix-on-write passive pass; 0: round-1 flux kernel + co
# 6: round-1 flux kernel + stability flag, convection decisions in lockstep form (co_decide_static) + region-wise averaging;
@pytest.mark.parametrize("mix", [6, 5, 4, 3, 2, 1, 0])
@pytest.mark.parametrize("nsteps", [5 * 40, 5 * 150])   # not the first steps: a uniform start is neutrally stable and
# the convection decisions there flip on the last bit (true of every non-strict variant)
def test_col_body_matches_oracle(nsteps, mix):
    _check_against_oracle(nsteps, mix)


# the lockstep form of the convection decisions (mix 6) visits the comparisons in another order than the reference's walk: more states,
# from the young ocean (most columns convect over many levels) to a two-year-old one
@pytest.mark.parametrize("nsteps", [5 * 80, 5 * 300, 5 * 480, 5 * 960])   # (from the uniform start every non-strict form flips decisions for ~20 ocean steps)
def test_lockstep_convection_decisions_match_oracle(nsteps):
    _check_against_oracle(nsteps, 6)


@pytest.mark.parametrize("nsteps,nrepeat", [(5 * 80, 12), (5 * 960, 8)])
def test_lockstep_convection_decisions_along_a_trajectory(nsteps, nrepeat):
    _check_against_oracle(nsteps, 6, nrepeat)


def _check_against_oracle(nsteps, mix, nrepeat=1):
    lib = _lib()
    oras = [Oracle("worjh2", maxk=K, maxl=L, nyear=96), Oracle("worjh2", maxk=K, maxl=L, nyear=96, diff1=2600.0, diff2=1.3e-5)]
    for o in oras:
        o.biogem_setup()
        o.run(nsteps)
        # BIOGEM rewrote the interior of ts/ts1; step_goldstein refreshes the periodic columns before tstepo
        # (goldstein.f90:176-187)
        ts, ts1 = (o.f(n).reshape(K + 2, J + 2, I + 2, L) for n in ("ts", "ts1"))
        ts1[:, :, 0, :] = ts[:, :, I, :]
        ts1[:, :, I + 1, :] = ts[:, :, 1, :]
    st = [_state(o) for o in oras]
    lanes = [m % 2 for m in range(MS)]          # even lanes: member A, odd lanes: member B

    def pack(name):
        return np.ascontiguousarray(np.stack([st[w][name] for w in lanes], axis=-1))

    ts_cur, u, tsflux, rho, cost = pack("ts"), pack("u"), pack("tsflux"), pack("rho"), pack("cost")
    ts_new = np.zeros_like(ts_cur)
    sst = np.zeros((2, J, I, MS))
    o0 = oras[0]
    k1 = o0.i("k1").astype(np.uint8).copy()
    k1i = o0.i("k1").reshape(J + 2, I + 2)
    cols = np.array([(i - 1) + I * (j - 1) for j in range(1, J + 1) for i in range(1, I + 1) if k1i[j, i] <= K], dtype=np.int32)
    par = lambda n: np.ascontiguousarray([oras[w].s(n) for w in lanes], dtype=np.float64)
    diff1, diff2 = par("diff1"), par("diff2")
    ec = np.ascontiguousarray(np.stack([par("ec%d" % q) for q in (1, 2, 3, 4)]))
    jm = np.ascontiguousarray(np.stack([o0.f(n)[:J + 2] for n in ("rc", "rc2", "cv", "cv2", "dsv", "rdsv", "rds")]))
    km = np.ascontiguousarray(np.stack([o0.f(n)[:K + 2] for n in ("dz", "dza", "rdz", "rdza", "ssmax")]))
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    wet = (k1i[1:J + 1, 1:I + 1][None, :, :] <= np.arange(1, K + 1)[:, None, None])
    nmixed = 0
    # nrepeat > 1: the tracer step repeated under the frozen flow and surface fluxes of the state (the oracle's tstepo leaves ts1 = ts
    # and the halo refreshed, :2410-2432): the decisions must stay the oracle's along a trajectory, not just for one step
    for rep in range(nrepeat):
        rc = lib.col_host_step(MS, dp(k1), dp(cols), len(cols), dp(ts_cur), dp(ts_new), dp(tsflux), dp(sst), dp(rho), dp(u),
                               dp(cost), dp(diff1), dp(diff2), dp(ec), dp(jm), dp(km), C.c_double(o0.s("dphi")),
                               C.c_double(o0.s("rdphi")), C.c_double(float(o0.f("dt")[K])), mix)
        assert rc == 0
        for w, o in enumerate(oras):
            before = o.f("cost").sum()
            o.call("tstepo")
            ts_ref, rho_ref, cost_ref = _after(o)
            for m in (w, w + 30):
                got = ts_new[..., m]
                scale = np.maximum(np.abs(ts_ref[wet]).max(axis=0), 1e-30)
                err = np.max(np.abs(got[wet] - ts_ref[wet]) / scale)
                erho = np.max(np.abs(rho[..., m][wet] - rho_ref[wet]) / np.abs(rho_ref[wet]).max())
                assert err < 1e-12 * (rep + 1), (rep, w, m, err)
                assert erho < 1e-12 * (rep + 1), (rep, w, m, erho)
                assert np.array_equal(cost[..., m], cost_ref), (rep, w, m)      # same mixing decisions, level for level
                assert np.array_equal(sst[0, :, :, m][wet[K - 1]], ts_ref[K - 1, :, :, 0][wet[K - 1]]) or \
                    np.max(np.abs(sst[0, :, :, m][wet[K - 1]] - ts_ref[K - 1, :, :, 0][wet[K - 1]])) < 1e-12 * (rep + 1)
            nmixed += int(cost_ref.sum() - before)
        ts_cur, ts_new = ts_new, ts_cur
    assert nmixed > 0          # the state exercises the convective adjustment


def test_lockstep_decisions_equal_the_walk_on_random_columns():
    """co_decide_static<WALK> (passes over a register-held column, in the reference's order of merges) against co_decide_core (the
    reference's walk on thread-private arrays) on 41 472 random columns per case: stable profiles with inversions of every length and
    size, exact ties between neighbouring levels (equal T and S: the reference mixes on `>=`), columns of every depth, passive tracers.
    Same partition in EVERY column (the convection counter `cost` equal), T / S / rho / passive tracers of the mixed boxes equal to
    rounding, every column inventory conserved, the result stable.  The form that merges every unstable run of a pass at once (mix 8)
    is held to the same on the moderate cases and may part from the walk where two unstable runs interact through the cubic equation
    of state (2 K of noise per level: reported)."""
    lib = _lib()
    o = Oracle("worjh2", maxk=K, maxl=L, nyear=96)
    dz = o.f("dz")[:K + 2].copy()
    km = np.ascontiguousarray(np.stack([o.f(n)[:K + 2] for n in ("dz", "dza", "rdz", "rdza", "ssmax")]))
    jm = np.ascontiguousarray(np.stack([o.f(n)[:J + 2] for n in ("rc", "rc2", "cv", "cv2", "dsv", "rdsv", "rds")]))
    ecv = np.array([o.s("ec%d" % q) for q in (1, 2, 3, 4)])
    ec = np.ascontiguousarray(np.repeat(ecv[:, None], MS, axis=1))
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    for case, (noise, tie_rate) in enumerate([(0.3, 0.0), (2.0, 0.0), (0.8, 0.3), (0.02, 0.6)]):
        rng = np.random.default_rng(100 + case)
        k1g = rng.integers(1, K + 1, size=(J + 2, I + 2)).astype(np.uint8)          # every cell wet, bottom level 1 .. K
        cols = np.arange(I * J, dtype=np.int32)
        lev = np.arange(K)[:, None, None, None]
        base_T = 2.0 + 18.0 * (lev / (K - 1.0)) ** 2                                  # warm on top: stable
        T = base_T + noise * rng.normal(size=(K, J, I, MS)) * (1.0 + 4.0 * (rng.random((1, J, I, MS)) < 0.3))
        S = 0.2 * rng.normal(size=(K, J, I, MS)) * noise
        tie = rng.random((K, J, I, MS)) < tie_rate                                     # level k takes the values of level k - 1
        for k in range(1, K):
            T[k] = np.where(tie[k], T[k - 1], T[k])
            S[k] = np.where(tie[k], S[k - 1], S[k])
        ts0 = np.zeros((K, J, I, L, MS))
        ts0[:, :, :, 0, :], ts0[:, :, :, 1, :] = T, S
        ts0[:, :, :, 2:, :] = rng.random((K, J, I, L - 2, MS)) + 0.5
        rho0 = ecv[0] * T + ecv[1] * S + ecv[2] * (T * T) + ecv[3] * (T * T * T)       # as eos() forms it (goldstein.f90:3048-3061)
        out = {}
        for mix in (7, 8, 9):
            ts, rho = ts0.copy(), np.ascontiguousarray(rho0.copy())
            cost, sst = np.zeros((J, I, MS)), np.zeros((2, J, I, MS))
            dummy3, dummyu = np.zeros((2, J, I, MS)), np.zeros((K, J, I, 3, MS))
            d1 = np.ones(MS)
            rc = lib.col_host_step(MS, dp(k1g), dp(cols), len(cols), dp(ts0), dp(ts), dp(dummy3), dp(sst), dp(rho), dp(dummyu), dp(cost),
                                   dp(d1), dp(d1), dp(ec), dp(jm), dp(km), C.c_double(o.s("dphi")), C.c_double(o.s("rdphi")),
                                   C.c_double(float(o.f("dt")[K])), mix)
            assert rc == 0
            out[mix] = (ts, rho, cost, sst)
        (tw, rw, cw, sw), (tl, rl, cl, sl) = out[7], out[9]
        k1 = k1g[1:J + 1, 1:I + 1].astype(int)
        wet = (np.arange(1, K + 1)[:, None, None] >= k1[None])[..., None, None]
        assert np.array_equal(cw, cl), "case %d: partitions differ in %d columns" % (case, int((cw != cl).sum()))
        nd = int((out[8][2] != cw).sum())
        print("case %d (noise %.2f K, ties %.0f %%): all-runs-at-once form differs from the walk in %d of %d columns" %
              (case, noise, 100 * tie_rate, nd, cw.size))
        assert nd <= 0.01 * cw.size
        assert cw.sum() > 0.2 * cw.size                                                # the adjustment is exercised
        scale = np.abs(ts0).reshape(-1, L, MS).max(axis=(0, 2))[None, None, None, :, None]
        err = np.abs(tw - tl) / scale
        assert err.max() <= 1e-14, (case, float(err.max()))
        assert np.abs(rw - rl).max() <= 1e-14 * np.abs(rw).max()
        assert np.abs(sw - sl).max() <= 1e-13
        w = dz[1:K + 1][:, None, None, None, None] * wet                               # thickness of the wet levels
        inv0 = (ts0 * w).sum(axis=0)
        for t in (tw, tl):
            assert np.abs((t * w).sum(axis=0) - inv0).max() <= 1e-12 * np.abs(inv0).max()
        # and the result is stable: no box is denser than (or as dense as) the box below it unless they are one box
        for t, r in ((tw, rw), (tl, rl)):
            for k in range(1, K):
                both = (k + 0 >= k1)[..., None]                                        # level k (index k - 1) and k + 1 wet
                same = (t[k, :, :, 0, :] == t[k - 1, :, :, 0, :]) & (t[k, :, :, 1, :] == t[k - 1, :, :, 1, :])
                assert (~both | same | (r[k] < r[k - 1])).all(), (case, k)
