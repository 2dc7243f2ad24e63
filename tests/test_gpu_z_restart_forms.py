"""ATCHEM's netCDF restart and the binary (ctrl_ncrst = .FALSE.) restarts of ATCHEM and BIOGEM moved between members of a device
ensemble (cgenie_b200/restart.py) -- run with -m gpu on a B200.  The file layer is covered on the CPU (tests/test_restart_nc.py)."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from cgenie_b200 import Ensemble, materialise

pytestmark = pytest.mark.gpu


def test_atchem_and_binary_restarts_through_device(built, tmp_path):
    """ATCHEM's netCDF restart (atchem_data_netCDF.f90:22-109, single precision) and the binary forms of ATCHEM's and BIOGEM's
    (atchem.f90:192-197, biogem.f90:2347-2355, double precision) written from one member of a device ensemble and read into
    another: the netCDF path carries float-rounded values, the binary path is bit-exact; the restarted member keeps running."""
    from cgenie_b200.restart import (ATM_TRACERS, OCN_TRACERS, SED_TRACERS, read_atchem_restart, read_biogem_restart_bin,
                                     write_atchem_restart, write_biogem_restart_bin)
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_ac_bg_36x36x16")
    k0 = np.array([2.0e-6, 1.7e-6])
    with Ensemble(str(job), n_members=2, perturb={"par_bio_k0_PO4": k0}) as e:
        e.run(5 * 40)
        I, J, K, L = e.maxi, e.maxj, e.maxk, e.maxl
        LA = len(ATM_TRACERS)
        src = e.get("atm", 1).reshape(J * I, LA).copy()
        p = write_atchem_restart(e, str(tmp_path / "rst" / "atchem_restart.nc"), member=1, year=0.4, run_id="test")
        with netcdf_file(p, "r", mmap=False) as f:
            assert len(f.variables) == 4 + LA and f.variables["atm_pCO2"].data.dtype == np.dtype(">f4")
            assert np.array_equal(f.variables["atm_pCO2"].data.ravel(), src[:, 2].astype(np.float32))
        assert read_atchem_restart(e, p, member=0) == [n for _, n, _ in ATM_TRACERS]
        got = e.get("atm", 0).reshape(J * I, LA)
        assert np.array_equal(got, src.astype(np.float32).astype(np.float64))
        assert np.array_equal(e.get("sfcatm1", 0).reshape(J * I, LA)[:, 2:], got[:, 2:])
        assert np.array_equal(e.get("atm", 1).reshape(J * I, LA), src)                       # the source member is untouched
        # binary forms: bit-exact
        pb = write_atchem_restart(e, str(tmp_path / "rst" / "atchem"), member=1, binary=True)
        assert read_atchem_restart(e, pb, member=0, binary=True) == [n for _, n, _ in ATM_TRACERS]
        assert np.array_equal(e.get("atm", 0).reshape(J * I, LA), src)
        src_ocn, src_part = e.get("ocn", 1).copy(), e.get("bio_part", 1).copy()
        pq = write_biogem_restart_bin(e, str(tmp_path / "rst" / "biogem"), member=1)
        assert read_biogem_restart_bin(e, pq, member=0) == [n for n, _ in OCN_TRACERS] + [n for n, _ in SED_TRACERS]
        assert np.array_equal(e.get("ocn", 0), src_ocn) and np.array_equal(e.get("bio_part", 0), src_part)
        k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
        wet = np.arange(1, K + 1)[:, None, None] >= k1[None]
        ts0 = e.get("ts", 0).reshape(K, J, I, L)
        assert np.array_equal(ts0[..., 0][wet], (src_ocn.reshape(K, J, I, L)[..., 0] - 273.15)[wet])
        e.run(5 * 4)
        assert int(e.health().sum()) == 0
