"""netCDF restart files in the reference's layout (csrc/cg_restart.cpp, cgenie_b200/restart.py): the C-ABI writers are
checked against an independent netCDF-3 reader (scipy.io.netcdf_file), the readers against files scipy writes, and the
layout against what outm_netcdf defines (goldstein_data.f90:153-300, embm_data.f90:83-200, gold_seaice_data.f90:100-230)."""
import ctypes as C

import numpy as np
import pytest
from scipy.io import netcdf_file

from cgenie_b200 import _lib
from cgenie_b200.restart import axes

I, J, K, Lt = 36, 36, 16, 16


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


@pytest.fixture(scope="module")
def state(built):
    rng = np.random.default_rng(3)
    k1 = np.full((J + 2, I + 2), 91, dtype=np.int32)          # (0:maxi+1, 0:maxj+1) column-major = [j][i]
    k1[1:J + 1, 1:I + 1] = rng.integers(1, K + 1, size=(J, I))
    k1[5:9, 7:12] = 91                                        # some land
    k1[20, 3] = 94
    ts = rng.normal(size=(K, J, I, Lt))                       # Fortran (maxl,maxi,maxj,maxk)
    u = rng.normal(size=(K, J, I, 3))
    tq = rng.normal(size=(J, I, 2))
    va = rng.uniform(size=(J, I, 2))
    tice, alb = rng.normal(size=(J, I)), rng.uniform(size=(J, I))
    s = np.concatenate([[0.0], np.linspace(-0.97, 0.97, J)])
    zro = -np.concatenate([[0.0], np.linspace(0.9, 0.01, K)])
    lon, lat, depth = axes(I, J, K, s, zro)
    date = np.array([2010, 7, 4, 123], dtype=np.int32)
    return dict(k1=k1, ts=ts, u=u, tq=tq, va=va, tice=tice, alb=alb, lon=lon, lat=lat, depth=depth, date=date)


def test_goldstein_restart_matches_reference_layout(state, tmp_path):
    L = _lib.load()
    p = str(tmp_path / "goldstein_restart_2010_07_04.nc")
    s = state
    assert L.cg_restart_goldstein_write(p.encode(), I, J, K, Lt, ip(s["k1"]), dp(s["lon"]), dp(s["lat"]), dp(s["depth"]),
                                        dp(s["ts"]), dp(s["u"]), None, None, None, ip(s["date"])) == 0
    with netcdf_file(p, "r", mmap=False) as f:
        assert f.version_byte == 1
        assert list(f.dimensions.items()) == [("nrecs", 1), ("longitude", I), ("latitude", J), ("depth", K)]
        assert list(f.variables) == ["longitude", "latitude", "depth", "ioffset", "iyear", "imonth", "iday", "temp",
                                     "salinity", "uvel", "vvel", "evap", "late", "sens"]
        assert f.variables["longitude"].units == b"degrees_east" and f.variables["latitude"].long_name == b"latitude"
        assert f.variables["longitude"].data.dtype == np.dtype(">f4") and f.variables["temp"].data.dtype == np.dtype(">f8")
        assert f.variables["iyear"].data.dtype == np.dtype(">i4") and f.variables["temp"].dimensions == ("depth", "latitude", "longitude")
        assert [int(f.variables[n].data[0]) for n in ("iyear", "imonth", "iday", "ioffset")] == [2010, 7, 4, 123]
        assert np.array_equal(f.variables["longitude"].data, s["lon"].astype(np.float32))
        assert np.array_equal(f.variables["depth"].data, s["depth"].astype(np.float32))
        ocean = (s["k1"][1:J + 1, 1:I + 1] <= K)[None, :, :]            # whole column written if any level is wet
        assert np.array_equal(f.variables["temp"].data, s["ts"][..., 0] * ocean)
        assert np.array_equal(f.variables["salinity"].data, s["ts"][..., 1] * ocean)
        assert np.array_equal(f.variables["uvel"].data, s["u"][..., 0]) and np.array_equal(f.variables["vvel"].data, s["u"][..., 1])
        assert not f.variables["evap"].data.any() and f.variables["sens"].data.shape == (J, I)
    # read back: T, S and u, v replaced, the other tracers and w untouched
    ts2, u2 = np.full_like(s["ts"], 7.0), np.full_like(s["u"], 7.0)
    date = np.zeros(4, dtype=np.int32)
    assert L.cg_restart_goldstein_read(p.encode(), I, J, K, Lt, dp(ts2), dp(u2), None, None, None, ip(date)) == 0
    assert np.array_equal(date, s["date"])
    assert np.array_equal(ts2[..., 0], s["ts"][..., 0] * ocean) and np.array_equal(u2[..., :2], s["u"][..., :2])
    assert (ts2[..., 2:] == 7.0).all() and (u2[..., 2] == 7.0).all()


def test_embm_and_seaice_restart_round_trip(state, tmp_path):
    L = _lib.load()
    s = state
    pe, ps = str(tmp_path / "embm.nc"), str(tmp_path / "sic.nc")
    assert L.cg_restart_embm_write(pe.encode(), I, J, dp(s["lon"]), dp(s["lat"]), dp(s["tq"]), ip(s["date"])) == 0
    assert L.cg_restart_seaice_write(ps.encode(), I, J, ip(s["k1"]), dp(s["lon"]), dp(s["lat"]), dp(s["va"]), dp(s["tice"]),
                                     dp(s["alb"]), ip(s["date"])) == 0
    with netcdf_file(pe, "r", mmap=False) as f:
        assert list(f.variables) == ["longitude", "latitude", "ioffset", "iyear", "imonth", "iday", "air_temp", "humidity"]
        assert np.array_equal(f.variables["air_temp"].data, s["tq"][..., 0]) and np.array_equal(f.variables["humidity"].data, s["tq"][..., 1])
        assert not hasattr(f.variables["longitude"], "units")            # only GOLDSTEIN's file carries axis attributes
    sea = s["k1"][1:J + 1, 1:I + 1] < 90
    with netcdf_file(ps, "r", mmap=False) as f:
        assert list(f.variables)[-4:] == ["sic_height", "sic_cover", "sic_temp", "sic_albedo"]
        assert np.array_equal(f.variables["sic_height"].data, s["va"][..., 0] * sea)
        assert np.array_equal(f.variables["sic_temp"].data, s["tice"])   # temperature and albedo are not masked
    tq, va, ti, al, date = np.empty(2 * I * J), np.empty(2 * I * J), np.empty(I * J), np.empty(I * J), np.zeros(4, dtype=np.int32)
    assert L.cg_restart_embm_read(pe.encode(), I, J, dp(tq), ip(date)) == 0 and np.array_equal(tq.reshape(J, I, 2), s["tq"])
    assert L.cg_restart_seaice_read(ps.encode(), I, J, dp(va), dp(ti), dp(al), ip(date)) == 0
    assert np.array_equal(va.reshape(J, I, 2), s["va"] * sea[..., None]) and np.array_equal(al.reshape(J, I), s["alb"])


def test_reads_files_written_by_another_netcdf3_writer(state, tmp_path):
    """A restart as netCDF-3 itself lays it out (scipy's writer: different variable order, an extra variable, a global
    attribute, 64-bit offsets) is read back bit for bit."""
    L = _lib.load()
    s = state
    for version in (1, 2):
        p = str(tmp_path / ("other%d.nc" % version))
        with netcdf_file(p, "w", version=version) as f:
            f.history = "written by scipy"
            for n, l in (("longitude", I), ("latitude", J), ("nrecs", 1)):
                f.createDimension(n, l)
            f.createVariable("humidity", "d", ("latitude", "longitude"))[:] = s["tq"][..., 1]
            f.createVariable("extra", "f", ("longitude",))[:] = 1.5
            f.createVariable("air_temp", "d", ("latitude", "longitude"))[:] = s["tq"][..., 0]
            for n, v in (("iday", 30), ("imonth", 12), ("iyear", 1999), ("ioffset", 0)):
                f.createVariable(n, "i", ("nrecs",))[:] = v
        tq, date = np.empty(2 * I * J), np.zeros(4, dtype=np.int32)
        assert L.cg_restart_embm_read(p.encode(), I, J, dp(tq), ip(date)) == 0
        assert np.array_equal(tq.reshape(J, I, 2), s["tq"]) and list(date) == [1999, 12, 30, 0]


def test_restart_errors_are_loud(state, tmp_path):
    L = _lib.load()
    tq, date = np.empty(2 * I * J), np.zeros(4, dtype=np.int32)
    assert L.cg_restart_embm_read(str(tmp_path / "nope.nc").encode(), I, J, dp(tq), ip(date)) == 2
    assert b"Missing file" in L.cg_restart_last_error()                  # the reference's message (embm_data.f90:30-33)
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"HDF5 is not what this reads")
    assert L.cg_restart_embm_read(str(bad).encode(), I, J, dp(tq), ip(date)) == 2 and b"not a netCDF classic" in L.cg_restart_last_error()
    s = state
    p = str(tmp_path / "small.nc")
    assert L.cg_restart_embm_write(p.encode(), I, J, dp(s["lon"]), dp(s["lat"]), dp(s["tq"]), ip(s["date"])) == 0
    assert L.cg_restart_embm_read(p.encode(), I + 1, J, dp(np.empty(2 * (I + 1) * J)), ip(date)) == 2
    assert b"wrong size" in L.cg_restart_last_error()
    assert L.cg_restart_seaice_read(p.encode(), I, J, dp(tq), dp(tq), dp(tq), ip(date)) == 2 and b"sic_height missing" in L.cg_restart_last_error()


def test_axes_from_host_constants(built, tmp_path):
    """The restart's coordinate variables from the library's own grid constants (cg_create only, no device): cGENIE's
    36x36x8 grid -- longitudes -255 .. 95 in steps of 10, sin(latitude) uniform, layer mid-depths 4234.5 .. 80.8 m."""
    from cgenie_b200 import materialise
    from test_host_init import HostOnly
    materialise(str(tmp_path), "eb_go_gs_36x36x8")
    h = HostOnly(str(tmp_path))
    try:
        lon, lat, depth = axes(36, 36, 8, h.const("s"), h.const("zro"))
    finally:
        h.close()
    assert np.allclose(lon, np.arange(-255.0, 100.0, 10.0), rtol=0, atol=1e-12)
    assert np.allclose(np.sin(np.radians(lat)), (np.arange(36) + 0.5) / 18.0 - 1.0, atol=1e-14)
    assert np.allclose(depth, [4234.5166, 3008.3391, 2099.7254, 1426.4307, 927.5105, 557.8040, 283.8467, 80.8407], atol=1e-3)


def test_biogem_restart_layout_and_round_trip(state, tmp_path):
    """BIOGEM's restart (biogem_data_netCDF.f90:24-142): FLOAT tracer variables, surface first, fill value on dry cells, the
    attributes sub_defvar writes; read back with the levels flipped and dry cells left alone."""
    from cgenie_b200.restart import OCN_TRACERS, SED_TRACERS, _strs, biogem_axes
    from cgenie_b200 import materialise
    from test_host_init import HostOnly
    L = _lib.load()
    job = tmp_path / "job"
    materialise(str(job), "eb_go_gs_ac_bg_36x36x16")
    h = HostOnly(str(job))
    try:
        k1 = h.iconst("k1")
        ax = [np.ascontiguousarray(a) for a in biogem_axes(I, J, K, h.const("s"), h.const("sv"), h.const("dz"), h.const("dza"))]
    finally:
        h.close()
    rng = np.random.default_rng(9)
    no, ns = len(OCN_TRACERS), len(SED_TRACERS)
    ocn = rng.uniform(1e-6, 3e-3, size=(K, J, I, no))
    part = rng.uniform(0, 1e-9, size=(K, J, I, ns))
    on, k_1 = _strs([n for n, _ in OCN_TRACERS]); ol, k_2 = _strs([l for _, l in OCN_TRACERS])
    sn, k_3 = _strs([n for n, _ in SED_TRACERS]); sl, k_4 = _strs([l for _, l in SED_TRACERS])
    p = str(tmp_path / "biogem_restart.nc")
    assert L.cg_restart_biogem_write(p.encode(), I, J, K, ip(k1), *[dp(a) for a in ax], no, on, ol, dp(ocn), ns, sn, sl, dp(part),
                                     123.0, b"run_x") == 0, L.cg_restart_last_error()
    k1ij = k1.reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = (np.arange(1, K + 1)[:, None, None] >= k1ij[None])             # [k][j][i], k = 1 deepest
    with netcdf_file(p, "r", mmap=False) as f:
        assert f.title == b"BIOGEM restart @ year 0000012" and f.Conventions == b"CF-1.0" and f.experiment_name == b"run_x"
        assert list(f.dimensions.items()) == [("lon", I), ("lat", J), ("lon_edges", I + 1), ("lat_edges", J + 1), ("zt", K), ("zt_edges", K + 1)]
        names = list(f.variables)
        assert names[:6] == ["lon", "lat", "lon_edges", "lat_edges", "zt", "zt_edges"]
        assert names[6:6 + no] == ["ocn_" + n for n, _ in OCN_TRACERS] and names[6 + no:] == ["bio_part_" + n for n, _ in SED_TRACERS]
        lon = f.variables["lon"]
        assert lon.axis == b"X" and lon.edges == b"lon_edges" and lon.units == b"degrees_east" and lon.standard_name == b"longitude"
        assert lon.missing_value == 9.9692099683868690e+36 and not hasattr(f.variables["lon_edges"], "axis")
        assert f.variables["zt"].units == b"cm" and f.variables["zt_edges"].units == b"m"
        assert np.allclose(f.variables["lon_edges"].data, np.arange(-260.0, 101.0, 10.0)) and f.variables["zt_edges"].data[0] == 0.0
        zt, zte = f.variables["zt"].data, f.variables["zt_edges"].data
        assert np.all(np.diff(zt) > 0) and np.all(zte[:-1] < zt) and np.all(zt < zte[1:]) and abs(zte[-1] - 5000.0) < 1e-9
        assert abs(zt[0] - 0.5 * zte[1]) < 1e-9            # the top level's mid-depth is half its thickness (biogem_data.f90:1105)
        v = f.variables["ocn_DIC"]
        assert v.data.dtype == np.dtype(">f4") and v.dimensions == ("zt", "lat", "lon")
        assert v.long_name == b"dissolved inorganic carbon (DIC)" and v.standard_name == b"Ocean tracer - DIC"
        assert v.missing_value == 9.9692099683868690e+36 and not hasattr(v, "units")
        want = np.where(wet, ocn[..., 2], 9.9692099683868690e+36)[::-1].astype(np.float32)
        assert np.array_equal(v.data, want)
        assert np.array_equal(f.variables["bio_part_CaCO3_frac2"].data, np.where(wet, part[..., 8], 9.9692099683868690e+36)[::-1].astype(np.float32))
    ocn2, part2 = np.full_like(ocn, -1.0), np.full_like(part, -1.0)
    fo, fs = np.zeros(no, dtype=np.int32), np.zeros(ns, dtype=np.int32)
    assert L.cg_restart_biogem_read(p.encode(), I, J, K, ip(k1), no, on, dp(ocn2), ip(fo), ns, sn, dp(part2), ip(fs)) == 0
    assert fo.all() and fs.all()
    w4 = np.broadcast_to(wet[..., None], ocn.shape)
    assert np.array_equal(ocn2[w4], ocn.astype(np.float32).astype(np.float64)[w4]) and (ocn2[~w4] == -1.0).all()
    assert np.array_equal(part2[..., 3][wet], part[..., 3].astype(np.float32).astype(np.float64)[wet])
    # a file with fewer tracers (another tracer selection): what it holds is taken, the rest is kept
    q = str(tmp_path / "partial.nc")
    with netcdf_file(q, "w") as f:
        f.createDimension("lon", I); f.createDimension("lat", J); f.createDimension("zt", K)
        f.createVariable("ocn_PO4", "f", ("zt", "lat", "lon"))[:] = 2.5e-6
    assert L.cg_restart_biogem_read(q.encode(), I, J, K, ip(k1), no, on, dp(ocn2), ip(fo), ns, sn, dp(part2), ip(fs)) == 0
    assert list(np.nonzero(fo)[0]) == [5] and not fs.any()
    assert np.all(ocn2[..., 5][wet] == np.float32(2.5e-6)) and np.array_equal(ocn2[..., 2][wet], ocn.astype(np.float32).astype(np.float64)[..., 2][wet])


def test_atchem_restart_layout_and_round_trip(built, tmp_path):
    """ATCHEM's netCDF restart (atchem_data_netCDF.f90:22-109): four dimensions, the axes sub_defvar describes, one FLOAT
    variable (lat, lon) per selected tracer with no mask; the reader replaces what the file holds (atchem_data.f90:152-163)."""
    from cgenie_b200.restart import ATM_TRACERS, _strs, atchem_axes
    L = _lib.load()
    na = len(ATM_TRACERS)
    rng = np.random.default_rng(21)
    atm = rng.uniform(1e-9, 3e-4, size=(J, I, na))                       # Fortran (n_atm, n_i, n_j)
    ax = [np.ascontiguousarray(a) for a in atchem_axes(I, J)]
    an, k_1 = _strs([n for _, n, _ in ATM_TRACERS]); al, k_2 = _strs([l for _, _, l in ATM_TRACERS])
    p = str(tmp_path / "atchem_restart.nc")
    assert L.cg_restart_atchem_write(p.encode(), I, J, *[dp(a) for a in ax], na, an, al, dp(atm), 10000.0, b"run_y") == 0, L.cg_restart_last_error()
    with netcdf_file(p, "r", mmap=False) as f:
        assert f.version_byte == 1
        assert f.title == b"ATCHEM restart @ year 0001000" and f.Conventions == b"CF-1.0" and f.experiment_name == b"run_y"
        assert list(f.dimensions.items()) == [("lon", I), ("lat", J), ("lon_edges", I + 1), ("lat_edges", J + 1)]
        assert list(f.variables) == ["lon", "lat", "lon_edges", "lat_edges"] + ["atm_" + n for _, n, _ in ATM_TRACERS]
        lat = f.variables["lat"]
        assert lat.axis == b"Y" and lat.edges == b"lat_edges" and lat.units == b"degrees_north" and lat.data.dtype == np.dtype(">f8")
        assert np.allclose(f.variables["lon"].data, np.arange(-255.0, 100.0, 10.0), rtol=0, atol=1e-12)
        assert np.allclose(f.variables["lon_edges"].data, np.arange(-260.0, 101.0, 10.0), rtol=0, atol=1e-12)
        assert np.allclose(np.sin(np.radians(lat.data)), (np.arange(J) + 0.5) / 18.0 - 1.0, atol=1e-14)
        late = f.variables["lat_edges"].data
        assert abs(late[0] + 90.0) < 1e-9 and abs(late[-1] - 90.0) < 1e-9 and np.all(late[:-1] < lat.data) and np.all(lat.data < late[1:])
        v = f.variables["atm_pCO2_13C"]
        assert v.data.dtype == np.dtype(">f4") and v.dimensions == ("lat", "lon")
        assert v.long_name == b"d13C CO2" and v.standard_name == b"Atmosphere tracer - pCO2_13C" and not hasattr(v, "units")
        assert v.missing_value == 9.9692099683868690e+36
        for l, (_, n, _) in enumerate(ATM_TRACERS):
            assert np.array_equal(f.variables["atm_" + n].data, atm[..., l].astype(np.float32)), n
    atm2 = np.full_like(atm, -1.0)
    fa = np.zeros(na, dtype=np.int32)
    assert L.cg_restart_atchem_read(p.encode(), I, J, na, an, dp(atm2), ip(fa)) == 0
    assert fa.all() and np.array_equal(atm2, atm.astype(np.float32).astype(np.float64))
    # a restart of a run with another selection: what it holds is taken, the rest kept
    q = str(tmp_path / "partial.nc")
    with netcdf_file(q, "w") as f:
        f.createDimension("lon", I); f.createDimension("lat", J)
        f.createVariable("atm_pCO2", "f", ("lat", "lon"))[:] = 278.0e-6
        f.createVariable("atm_pCH4", "f", ("lat", "lon"))[:] = 7.0e-7
    assert L.cg_restart_atchem_read(q.encode(), I, J, na, an, dp(atm2), ip(fa)) == 0
    assert list(np.nonzero(fa)[0]) == [2] and np.all(atm2[..., 2] == np.float32(278.0e-6))
    assert np.array_equal(atm2[..., 5], atm[..., 5].astype(np.float32).astype(np.float64))
    # wrong grid and missing file are errors with a message
    assert L.cg_restart_atchem_read(p.encode(), I, J + 1, na, an, dp(atm2), ip(fa)) != 0 and b"wrong size" in L.cg_restart_last_error()
    assert L.cg_restart_atchem_read(str(tmp_path / "none.nc").encode(), I, J, na, an, dp(atm2), ip(fa)) != 0


def _fortran_record(path):
    """Independent reader of a single-record gfortran unformatted sequential file."""
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[:4], "<i4")[0])
    assert len(raw) == n + 8 and int(np.frombuffer(raw[-4:], "<i4")[0]) == n
    return raw[4:-4]


def test_binary_restarts_are_fortran_records(built, tmp_path):
    """ctrl_ncrst = .FALSE.: atchem.f90:192-197 and biogem.f90:2347-2355 write ONE unformatted record (count, global tracer
    indices, the array sections in selection order; INTEGER*4 / REAL*8); checked byte for byte against numpy, read back exactly,
    and read by a caller with another selection (biogem_data.f90:540-547 matches on the indices the record carries)."""
    from cgenie_b200.restart import ATM_TRACERS, OCN_IDS, SED_IDS
    L = _lib.load()
    rng = np.random.default_rng(22)
    na, no, ns = len(ATM_TRACERS), len(OCN_IDS), len(SED_IDS)
    ia = np.array([a for a, _, _ in ATM_TRACERS], dtype=np.int32)
    io, isd = np.array(OCN_IDS, dtype=np.int32), np.array(SED_IDS, dtype=np.int32)
    atm = rng.normal(size=(J, I, na))
    ocn, part = rng.normal(size=(K, J, I, no)), rng.normal(size=(K, J, I, ns))
    pa, pb = str(tmp_path / "atchem"), str(tmp_path / "biogem")
    assert L.cg_restart_atchem_write_bin(pa.encode(), I, J, na, ip(ia), dp(atm)) == 0
    assert L.cg_restart_biogem_write_bin(pb.encode(), I, J, K, no, ip(io), dp(ocn), ns, ip(isd), dp(part)) == 0
    want = np.int32(na).tobytes() + ia.tobytes() + b"".join(np.ascontiguousarray(atm[..., l]).tobytes() for l in range(na))
    assert _fortran_record(pa) == want
    want = (np.int32(no).tobytes() + io.tobytes() + b"".join(np.ascontiguousarray(ocn[..., l]).tobytes() for l in range(no)) +
            np.int32(ns).tobytes() + isd.tobytes() + b"".join(np.ascontiguousarray(part[..., l]).tobytes() for l in range(ns)))
    assert _fortran_record(pb) == want
    atm2, ocn2, part2 = np.zeros_like(atm), np.zeros_like(ocn), np.zeros_like(part)
    fa, fo, fs = np.zeros(na, dtype=np.int32), np.zeros(no, dtype=np.int32), np.zeros(ns, dtype=np.int32)
    assert L.cg_restart_atchem_read_bin(pa.encode(), I, J, na, ip(ia), dp(atm2), ip(fa)) == 0
    assert L.cg_restart_biogem_read_bin(pb.encode(), I, J, K, no, ip(io), dp(ocn2), ip(fo), ns, ip(isd), dp(part2), ip(fs)) == 0
    assert fa.all() and fo.all() and fs.all()
    assert np.array_equal(atm2, atm) and np.array_equal(ocn2, ocn) and np.array_equal(part2, part)      # bit-exact, unlike netCDF
    # a caller that selected fewer tracers, in another order, plus one the file does not hold
    sel = np.array([12, 3, 99, 1], dtype=np.int32)                     # ALK, DIC, (absent), temp
    o3 = np.full((K, J, I, 4), -7.0)
    f3 = np.zeros(4, dtype=np.int32)
    assert L.cg_restart_biogem_read_bin(pb.encode(), I, J, K, 4, ip(sel), dp(o3), ip(f3), 0, None, None, None) == 0
    assert list(f3) == [1, 1, 0, 1]
    assert np.array_equal(o3[..., 0], ocn[..., 7]) and np.array_equal(o3[..., 1], ocn[..., 2]) and np.array_equal(o3[..., 3], ocn[..., 0])
    assert (o3[..., 2] == -7.0).all()
    # wrong grid: the record's length does not fit
    assert L.cg_restart_biogem_read_bin(pb.encode(), I, J, K - 1, no, ip(io), dp(ocn2), ip(fo), ns, ip(isd), dp(part2), ip(fs)) != 0
    assert b"BIOGEM restart" in L.cg_restart_last_error()
    assert L.cg_restart_atchem_read_bin(pb.encode(), I, J, na, ip(ia), dp(atm2), ip(fa)) != 0
    open(str(tmp_path / "trunc"), "wb").write(open(pa, "rb").read()[:-9])
    assert L.cg_restart_atchem_read_bin(str(tmp_path / "trunc").encode(), I, J, na, ip(ia), dp(atm2), ip(fa)) != 0


@pytest.mark.parametrize("name, date", [("main_restart_0.nc", (2000, 1, 2, 0)), ("main_fluxes_0_date.nc", (1999, 12, 30, 0))])
def test_codec_against_the_netcdf_files_the_reference_ships(built, tmp_path, name, date):
    """data/main/main_restart_0.nc and main_fluxes_0_date.nc are the two files in the reference tree that the netCDF library itself
    wrote (hex copies under tests/golden/, tools/make_golden.py ref_date_files): the date block every module's restart starts
    with.  The reader must take the date out of them and the writer must reproduce them BYTE FOR BYTE -- header layout, name
    padding, begin offsets, big-endian INT data of the codec are the library's."""
    import binascii
    import os
    L = _lib.load()
    ref = binascii.unhexlify(open(os.path.join(os.path.dirname(__file__), "golden", "ref_" + name + ".hex")).read().strip())
    p = str(tmp_path / name)
    open(p, "wb").write(ref)
    got = np.zeros(4, dtype=np.int32)
    assert L.cg_restart_date_read(p.encode(), ip(got)) == 0, L.cg_restart_last_error()
    assert tuple(int(x) for x in got) == date                       # {iyear, imonth, iday, ioffset}
    q = str(tmp_path / ("mine_" + name))
    assert L.cg_restart_date_write(q.encode(), ip(np.array(date, dtype=np.int32))) == 0, L.cg_restart_last_error()
    assert open(q, "rb").read() == ref
    # and the module restarts' readers accept the date block of such a file's layout (same variable names and types)
    with netcdf_file(q, "r", mmap=False) as f:
        assert list(f.variables) == ["ioffset", "iyear", "imonth", "iday"] and list(f.dimensions.items()) == [("nrecs", 1)]
