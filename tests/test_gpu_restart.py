"""netCDF restarts of one ensemble member written from and read into the device state (cgenie_b200/restart.py) -- run
with -m gpu on a B200.  The file layer itself is covered on the CPU (tests/test_restart_nc.py)."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.restart import read_restart, write_restart

pytestmark = pytest.mark.gpu


def test_restart_round_trip_through_device(built, tmp_path):
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8")
    M = 3
    diff1 = np.array([2000.0, 1700.0, 2300.0])
    names = ("ts", "u", "tq", "tq1", "varice", "varice1", "tice", "albice")
    with Ensemble(str(job), n_members=M, perturb={"diff1": diff1}) as e:
        e.run(5 * 40)
        paths = write_restart(e, str(tmp_path / "rst"), member=1, date=[2003, 4, 12, 0])
        assert sorted(p.rsplit("/", 1)[1] for p in paths.values()) == [
            "embm_restart_2003_04_12.nc", "goldsic_restart_2003_04_12.nc", "goldstein_restart_2003_04_12.nc"]
        want = {n: e.get(n, 1).copy() for n in names}
        I, J, K, L = e.maxi, e.maxj, e.maxk, e.maxl
        with netcdf_file(paths["goldstein"], "r", mmap=False) as f:
            ts = want["ts"].reshape(K, J, I, L)
            k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
            assert np.array_equal(f.variables["temp"].data, ts[..., 0] * (k1 <= K)[None])
            assert f.variables["temp"].data.any() and f.variables["depth"].data[0] > f.variables["depth"].data[-1] > 0
        # member 2 takes member 1's restart; member 0 is left alone
        before0 = {n: e.get(n, 0).copy() for n in names}
        date = read_restart(e, paths, member=2)
        assert list(date) == [2003, 4, 12, 0]
        got = {n: e.get(n, 2) for n in names}
        ocean = np.broadcast_to((k1 <= K)[None, :, :, None], (K, J, I, L))
        assert np.array_equal(got["ts"].reshape(K, J, I, L)[ocean], want["ts"].reshape(K, J, I, L)[ocean])
        assert np.array_equal(got["u"].reshape(K, J, I, 3)[..., :2], want["u"].reshape(K, J, I, 3)[..., :2])
        for n in ("tq", "tice", "albice"):
            assert np.array_equal(got[n], want[n]), n
        assert np.array_equal(got["tq1"], want["tq"]) and np.array_equal(got["varice1"], got["varice"])
        sea = (k1 < 90)[..., None]
        assert np.array_equal(got["varice"].reshape(J, I, 2), want["varice"].reshape(J, I, 2) * sea)
        for n in names:
            assert np.array_equal(e.get(n, 0), before0[n]), n
        e.run(5 * 4)                      # the restarted member keeps running
        assert int(e.health().sum()) == 0


def test_biogem_restart_through_device(built, tmp_path):
    """BIOGEM's netCDF restart (single precision, as the reference stores it) of one member written from the device and
    read into another member: wet cells carry the float-rounded values, the second member keeps running."""
    from cgenie_b200.restart import OCN_TRACERS, SED_TRACERS, read_biogem_restart, write_biogem_restart
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_ac_bg_36x36x16")
    k0 = np.array([2.0e-6, 1.7e-6])
    with Ensemble(str(job), n_members=2, perturb={"par_bio_k0_PO4": k0}) as e:
        e.run(5 * 40)
        I, J, K, L = e.maxi, e.maxj, e.maxk, e.maxl
        LS = len(SED_TRACERS)
        p = write_biogem_restart(e, str(tmp_path / "rst" / "biogem_restart.nc"), member=1, year=0.4, run_id="test")
        with netcdf_file(p, "r", mmap=False) as f:
            assert len(f.variables) == 6 + L + LS and f.variables["ocn_temp"].data.dtype == np.dtype(">f4")
            sst = f.variables["ocn_temp"].data[0]
            assert 270.0 < sst[sst < 1e30].min() and sst[sst < 1e30].max() < 310.0      # BIOGEM keeps temperature in K
        src_ocn = e.get("ocn", 1).reshape(K, J, I, L).copy()
        src_part = e.get("bio_part", 1).reshape(K, J, I, LS).copy()
        found = read_biogem_restart(e, p, member=0)
        assert found == [n for n, _ in OCN_TRACERS] + [n for n, _ in SED_TRACERS]
        k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
        wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
        got_ocn = e.get("ocn", 0).reshape(K, J, I, L)
        got_part = e.get("bio_part", 0).reshape(K, J, I, LS)
        for l in range(L):
            assert np.array_equal(got_ocn[..., l][wet], src_ocn[..., l].astype(np.float32).astype(np.float64)[wet]), OCN_TRACERS[l]
        for l in range(LS):
            assert np.array_equal(got_part[..., l][wet], src_part[..., l].astype(np.float32).astype(np.float64)[wet]), SED_TRACERS[l]
        assert np.array_equal(e.get("ocn", 1).reshape(K, J, I, L), src_ocn)              # the source member is untouched
        # GOLDSTEIN's ts of the restarted member: T, S from ocn (ctrl_force_GOLDSTEInTS), the other tracers salinity-normalised
        ts0 = e.get("ts", 0).reshape(K, J, I, L)
        V = e.const("bg_V").reshape(K, J, I)
        mean_S = (got_ocn[..., 1] * V)[wet].sum() / V[wet].sum()
        assert np.array_equal(ts0[..., 0][wet], (got_ocn[..., 0] - 273.15)[wet])
        assert np.allclose(ts0[..., 1][wet], got_ocn[..., 1][wet] - 34.9, rtol=0, atol=1e-12)
        for l in (2, 5, 15):
            assert np.allclose(ts0[..., l][wet], got_ocn[..., l][wet] * mean_S / got_ocn[..., 1][wet], rtol=1e-12, atol=0)
        e.run(5 * 4)
        assert int(e.health().sum()) == 0


def test_run_continued_from_restart_matches_uninterrupted(built, tmp_path):
    """A fresh handle started from the restart files of a running member (inm_netcdf: T, S, u with u1 = u, rho from eos,
    tq1 = tq, varice1 = varice) continues bit for bit like the member it was written from: GOLDSTEIN's restart holds
    doubles, and every other prognostic quantity is rebuilt from the restored ones within the first cycle."""
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_36x36x8")
    diff1 = np.array([2000.0, 1700.0])
    names = ("ts", "u", "u1", "rho", "tq", "varice", "tice")
    with Ensemble(str(job), n_members=2, perturb={"diff1": diff1}) as e:
        e.set_tracer_variant("strict")
        e.run(5 * 40)
        paths = write_restart(e, str(tmp_path / "rst"), member=1, date=[2003, 4, 12, 0])
        e.run(5 * 20)
        want = {n: e.get(n, 1).copy() for n in names}
        K, J, I, L = e.maxk, e.maxj, e.maxi, e.maxl
        k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    with Ensemble(str(job), n_members=1, perturb={"diff1": diff1[1:]}) as f:
        f.set_tracer_variant("strict")
        read_restart(f, paths, member=0)
        f.set_koverall(5 * 40)
        f.run(5 * 20)
        wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
        for n in names:
            got, ref = f.get(n, 0), want[n]
            if n in ("ts", "rho", "u", "u1"):
                c = got.size // (K * J * I)
                got, ref = got.reshape(K, J, I, c)[wet], ref.reshape(K, J, I, c)[wet]
            assert np.array_equal(got, ref), (n, float(np.abs(got - ref).max()))
        assert int(f.health().sum()) == 0
