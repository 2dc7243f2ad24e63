"""A C host program (tests/c_driver/genie_driver.c, built with gcc against include/cgenie_b200.h and libcgenie_b200.so) drives
the library in genie.f90's exact call order with Fortran-shaped column-major host arrays, as the ISO_C_BINDING shims of
fortran/ would from a linked genie.exe -- no Python between host program and library.  Its output is compared with the oracle
advanced in lock step.  Run with -m gpu on a B200."""
import os
import subprocess

import numpy as np
import pytest

from cgenie_b200 import materialise
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cdrv") / "genie_driver")
    lib = os.path.join(ROOT, "cgenie_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "c_driver", "genie_driver.c"),
                           "-L" + lib, "-lcgenie_b200", "-lm", "-Wl,-rpath," + lib])
    return exe


def read_out(path, biogem):
    raw = np.fromfile(path, dtype=np.float64)
    L, I, J, K, n_out = (int(x) for x in raw[:5])
    ij, ijk = I * J, I * J * K
    p = 5
    outs = []
    for _ in range(n_out):
        rec = {}
        for n in ("tstar_ocn", "sstar_ocn", "tstar_atm", "qstar_atm", "hght_sic", "frac_sic", "latent_ocn"):
            rec[n] = raw[p:p + ij]; p += ij
        rec["rho"] = raw[p:p + ijk]; p += ijk
        rec["test_energy_ocean"], rec["test_water_ocean"] = raw[p], raw[p + 1]; p += 2
        outs.append(rec)
    fin = {"ts": raw[p:p + ijk * L]}
    p += ijk * L
    if biogem:
        fin["ts1"] = raw[p:p + ijk * L]; p += ijk * L
        fin["ocn"] = raw[p:p + ijk * L]; p += ijk * L
        fin["atm"] = raw[p:]
    else:
        assert p == raw.size
    return (L, I, J, K), outs, fin


@pytest.mark.parametrize("cfg,okw,nk,every,tol", [
    ("eb_go_gs_36x36x8", dict(world="worbe2", maxk=8, maxl=2, nyear=100), 60, 3, 1e-9),
    ("eb_go_gs_ac_bg_36x36x16", dict(world="worjh2", maxk=16, maxl=16, nyear=96), 40, 2, 1e-9)])
def test_c_host_in_genie_order(driver, tmp_path, cfg, okw, nk, every, tol):
    job = tmp_path / "job"
    materialise(str(job), cfg)
    out = str(tmp_path / "out.bin")
    res = subprocess.run([driver, str(job), str(nk), str(every), out, "3"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    biogem = "bg" in cfg
    (L, I, J, K), outs, fin = read_out(out, biogem)
    assert len(outs) == (nk // 5) // every and L == okw["maxl"] and K == okw["maxk"]
    o = Oracle(**okw)
    if biogem:
        o.biogem_setup()

    def close(a, b, what):
        # the library runs its default 'strict' tracer variant: everything but surflux's exp / log / pow (1.5e-14 per call, CUDA
        # libm vs glibc) and BIOGEM's carbonate chemistry is bit-identical to the oracle; per cell, floored at 1e-3 of the field
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * max(np.abs(b).max(), 1e-300))
        assert err.max() <= tol, (what, float(err.max()))

    rec = iter(outs)
    for n in range(1, nk // 5 + 1):
        o.run(5)
        if n % every:
            continue
        r = next(rec)
        for name in ("tstar_ocn", "sstar_ocn", "tstar_atm", "qstar_atm", "hght_sic", "frac_sic", "latent_ocn"):
            close(r[name], o.f(name), "%s at ocean step %d" % (name, n))
        rho = o.f("rho").reshape(K + 1, J + 2, I + 2)[1:, 1:J + 1, 1:I + 1].ravel()
        close(r["rho"], rho, "go_rho at ocean step %d" % n)
        for name in ("test_energy_ocean", "test_water_ocean"):
            ref = o.s(name)
            assert abs(r[name] - ref) <= 1e-6 * max(abs(ref), 1.0), (name, r[name], ref)
    ts = o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :]
    k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = np.broadcast_to((np.arange(1, K + 1)[:, None, None] >= k1[None])[..., None], ts.shape)
    # the end state after nk / 5 ocean steps from the uniform, neutrally stable initial state: a convective adjustment may flip on
    # the last bit of surflux's libm differences (tests/test_gpu_biogem.py::test_biogem_model_steps holds 10 BIOGEM steps to 1e-6)
    tol = max(tol, 1e-6)
    close(fin["ts"].reshape(ts.shape)[wet], ts[wet], "final go_ts")
    if biogem:
        assert np.array_equal(fin["ts1"], fin["ts"])                     # go_ts1 = go_ts (biogem.f90:2055-2056)
        close(fin["ocn"].reshape(ts.shape)[wet], o.f("ocn").reshape(ts.shape)[wet], "ocn")
        close(fin["atm"], o.f("atm"), "atm")
