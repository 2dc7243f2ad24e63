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


def _check_against_oracle(nsteps, mix):
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
    rc = lib.col_host_step(MS, dp(k1), dp(cols), len(cols), dp(ts_cur), dp(ts_new), dp(tsflux), dp(sst), dp(rho), dp(u),
                           dp(cost), dp(diff1), dp(diff2), dp(ec), dp(jm), dp(km), C.c_double(o0.s("dphi")),
                           C.c_double(o0.s("rdphi")), C.c_double(float(o0.f("dt")[K])), mix)
    assert rc == 0
    wet = (k1i[1:J + 1, 1:I + 1][None, :, :] <= np.arange(1, K + 1)[:, None, None])
    nmixed = 0
    for w, o in enumerate(oras):
        o.call("tstepo")
        ts_ref, rho_ref, cost_ref = _after(o)
        for m in (w, w + 30):
            got = ts_new[..., m]
            scale = np.maximum(np.abs(ts_ref[wet]).max(axis=0), 1e-30)
            err = np.max(np.abs(got[wet] - ts_ref[wet]) / scale)
            erho = np.max(np.abs(rho[..., m][wet] - rho_ref[wet]) / np.abs(rho_ref[wet]).max())
            assert err < 1e-12, (w, m, err)
            assert erho < 1e-12, (w, m, erho)
            assert np.array_equal(cost[..., m], cost_ref)       # same mixing decisions, level for level
            assert np.array_equal(sst[0, :, :, m][wet[K - 1]], ts_ref[K - 1, :, :, 0][wet[K - 1]]) or \
                np.max(np.abs(sst[0, :, :, m][wet[K - 1]] - ts_ref[K - 1, :, :, 0][wet[K - 1]])) < 1e-12
        nmixed += int((cost_ref - st[w]["cost"]).sum())
    assert nmixed > 0          # the state exercises the convective adjustment
