"""netCDF restart files of GOLDSTEIN, the EMBM and the sea-ice model in the reference's layout, for hosts without the
Fortran model.  The file format and the per-module layouts live in the C-ABI library
(csrc/cg_restart.cpp: cg_restart_*_write / _read on plain arrays, no device involved); this module only moves one
member's state between an `Ensemble` and those calls, as outm_netcdf / inm_netcdf do between the module arrays and the
netCDF library (goldstein_data.f90:11-300, embm_data.f90:11-200, gold_seaice_data.f90:11-230).

BIOGEM / ATCHEM keep their own restart files in the reference (biogem_data_netCDF.f90:24-147); those are not written here:
the biogeochemical tracers of `ts` beyond T and S are not part of GOLDSTEIN's restart either (goldstein_data.f90:209-213).
"""
import ctypes as C
import os

import numpy as np

from . import _lib

DSC = 5.0e3   # goldstein_lib.f90:49


class RestartError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _ck(rc):
    if rc:
        raise RestartError(_lib.load().cg_restart_last_error().decode())


def axes(maxi, maxj, maxk, s, zro):
    """nclon1, nclat1 (goldstein.f90:1843-1885, igrid 0) and the restart's depth axis depths1(k) = ncdepth(maxk-k+1) =
    |dsc zro(k)| (goldstein_data.f90:194-198, goldstein.f90:1920-1923); s = sin(latitude) (0:maxj), zro (0:maxk)."""
    i = np.arange(1, maxi + 1, dtype=np.float64)
    lon = 360.0 * (i - 0.5) / float(maxi) + (-260.0)
    lat = np.arcsin(np.asarray(s, dtype=np.float64)[1:maxj + 1]) * 180.0 / np.pi
    depth = np.abs(DSC * np.asarray(zro, dtype=np.float64)[1:maxk + 1])
    return lon, lat, depth


def default_date(e):
    """iyear_rest, imonth_rest, day, ioffset_rest of a run without restart input (goldstein.f90:1729-1732)."""
    return np.array([2000, 1, int(round(360.0 / e.nyear)), 0], dtype=np.int32)


def write_restart(e, outdir, member=0, date=None):
    """goldstein_restart_*, embm_restart_*, goldsic_restart_* .nc of one member, named yyyy_mm_dd like the reference."""
    L = _lib.load()
    I, J, K, Lt = e.maxi, e.maxj, e.maxk, e.maxl
    date = default_date(e) if date is None else np.ascontiguousarray(date, dtype=np.int32)
    k1 = np.ascontiguousarray(e.iconst("k1"), dtype=np.int32)
    lon, lat, depth = axes(I, J, K, e.const("s"), e.const("zro"))
    tag = "%d_%02d_%02d.nc" % (date[0], date[1], date[2])
    os.makedirs(outdir, exist_ok=True)
    paths = {m: os.path.join(outdir, "%s_restart_%s" % (m, tag)) for m in ("goldstein", "embm", "goldsic")}
    ts = np.ascontiguousarray(e.get("ts", member), dtype=np.float64)
    u = np.ascontiguousarray(e.get("u", member), dtype=np.float64)
    _ck(L.cg_restart_goldstein_write(paths["goldstein"].encode(), I, J, K, Lt, _ip(k1), _dp(lon), _dp(lat), _dp(depth),
                                     _dp(ts), _dp(u), None, None, None, _ip(date)))
    tq = np.ascontiguousarray(e.get("tq", member), dtype=np.float64)
    _ck(L.cg_restart_embm_write(paths["embm"].encode(), I, J, _dp(lon), _dp(lat), _dp(tq), _ip(date)))
    va = np.ascontiguousarray(e.get("varice", member), dtype=np.float64)
    tice = np.ascontiguousarray(e.get("tice", member), dtype=np.float64)
    alb = np.ascontiguousarray(e.get("albice", member), dtype=np.float64)
    _ck(L.cg_restart_seaice_write(paths["goldsic"].encode(), I, J, _ip(k1), _dp(lon), _dp(lat), _dp(va), _dp(tice), _dp(alb),
                                  _ip(date)))
    return paths


def read_restart(e, paths, member=0):
    """inm_netcdf of the three modules: T, S and the horizontal velocities (u1 = u), air temperature and humidity
    (tq1 = tq), sea-ice height, cover (varice1 = varice), temperature and albedo of one member.  Returns the files' date."""
    L = _lib.load()
    I, J, K, Lt = e.maxi, e.maxj, e.maxk, e.maxl
    date = np.zeros(4, dtype=np.int32)
    ts = np.ascontiguousarray(e.get("ts", member), dtype=np.float64)
    u = np.ascontiguousarray(e.get("u", member), dtype=np.float64)
    _ck(L.cg_restart_goldstein_read(paths["goldstein"].encode(), I, J, K, Lt, _dp(ts), _dp(u), None, None, None, _ip(date)))
    e.put("ts", ts, member)
    e.put("u", u, member)
    tq = np.empty(2 * I * J)
    _ck(L.cg_restart_embm_read(paths["embm"].encode(), I, J, _dp(tq), _ip(date)))
    e.put("tq", tq, member)
    e.put("tq1", tq, member)
    va, tice, alb = np.empty(2 * I * J), np.empty(I * J), np.empty(I * J)
    _ck(L.cg_restart_seaice_read(paths["goldsic"].encode(), I, J, _dp(va), _dp(tice), _dp(alb), _ip(date)))
    e.put("varice", va, member)
    e.put("varice1", va, member)
    e.put("tice", tice, member)
    e.put("albice", alb, member)
    return date
