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
