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


# ------------------------------------------------------------------ BIOGEM (frozen tracer selection)
# string_ocn / string_longname_ocn and string_sed / string_longname_sed of the selected tracers, in conv_iselected order
# (data/main/tracer_define.ocn, tracer_define.sed: columns 1 and 5)
OCN_TRACERS = [("temp", "temperature"), ("sal", "salinity"), ("DIC", "dissolved inorganic carbon (DIC)"),
               ("DIC_13C", "d13C of DIC"), ("DIC_14C", "d14C of DIC"), ("PO4", "dissolved phosphate (PO4)"),
               ("O2", "dissolved oxygen (O2)"), ("ALK", "alkalinity (ALK)"),
               ("DOM_C", "dissolved organic matter (DOM); carbon"), ("DOM_C_13C", "d13C of DOM-C"),
               ("DOM_C_14C", "d14C of DOM-C"), ("DOM_P", "dissolved organic matter; phosphorous"),
               ("Ca", "dissolved calcium (Ca)"), ("CFC11", "dissolved CFC-11"), ("CFC12", "dissolved CFC-12"),
               ("Mg", "dissolved Magnesium (Mg)")]
SED_TRACERS = [("POC", "particulate organic carbon (POC)"), ("POC_13C", "d13C of POC"), ("POC_14C", "d14C of POC"),
               ("POP", "particulate organic phosphate (POP)"), ("CaCO3", "calcium carbonate (CaCO3)"),
               ("CaCO3_13C", "d13C of CaCO3"), ("CaCO3_14C", "d14C of CaCO3"), ("POC_frac2", "n/a"), ("CaCO3_frac2", "n/a")]


def _strs(items):
    arr = (C.c_char_p * len(items))(*[x.encode() for x in items])
    return C.cast(arr, C.POINTER(C.c_char_p)), arr


def biogem_axes(n_i, n_j, n_k, s, sv, dz, dza):
    """The coordinate variables of BIOGEM's files: phys_ocn lon / lat / Dmid, their edges through edge_maker
    (biogem_data.f90:1115-1123, gem_netcdf.f90:877-913, biogem_data_netCDF.f90:107-119).  s, sv (0:n_j), dz, dza (0:n_k)
    are GOLDSTEIN's grid arrays; par_grid_lon_offset = -260."""
    off = -260.0
    lon = np.array([(360.0 / n_i) * (float(i) - 0.5) + off for i in range(1, n_i + 1)])
    lone = np.array([(360.0 / n_i) * float(i) + off for i in range(1, n_i + 1)])
    lon_e = np.concatenate([[lone[0] - 360.0 / n_i], lone])
    rad = 180.0 / np.pi
    lat = np.array([rad * np.arcsin(s[j]) for j in range(1, n_j + 1)])
    latn = np.array([rad * np.arcsin(sv[j]) for j in range(1, n_j + 1)])
    dlat1 = rad * (np.arcsin(sv[1]) - np.arcsin(sv[0]))
    lat_e = np.concatenate([[latn[0] - dlat1], latn])

    dza = np.array(dza, dtype=np.float64)
    dza[n_k] = dz[n_k] / 2.0     # loc_grid_dza(n_k) = loc_grid_dz(n_k)/2.0, biogem_data.f90:1105

    def tail_sum(a, k):          # SUM(dsc * a(k:n_k)), added in index order
        t = 0.0
        for q in range(k, n_k + 1):
            t = t + DSC * a[q]
        return t
    dmid = [tail_sum(dza, k) for k in range(1, n_k + 1)]
    dbot = [tail_sum(dz, k) for k in range(1, n_k + 1)]
    zt = np.array(dmid[::-1])
    zt_e = np.concatenate([[0.0], dbot[::-1]])       # loc_zt_e(0) = 0.0 overrides edge_maker's first edge
    return lon, lat, lon_e, lat_e, zt, zt_e


def write_biogem_restart(e, path, member=0, year=0.0, run_id=""):
    """BIOGEM's netCDF restart of one member (ocean tracers and particulates as FLOAT variables, as the reference stores them)."""
    L = _lib.load()
    I, J, K = e.maxi, e.maxj, e.maxk
    if e.maxl != len(OCN_TRACERS):
        raise RestartError("BIOGEM restart: the job's tracer selection is not the frozen one")
    k1 = np.ascontiguousarray(e.iconst("k1"), dtype=np.int32)
    ax = [np.ascontiguousarray(a, dtype=np.float64) for a in
          biogem_axes(I, J, K, e.const("s"), e.const("sv"), e.const("dz"), e.const("dza"))]
    ocn = np.ascontiguousarray(e.get("ocn", member), dtype=np.float64)
    part = np.ascontiguousarray(e.get("bio_part", member), dtype=np.float64)
    on, keep1 = _strs([n for n, _ in OCN_TRACERS])
    ol, keep2 = _strs([l for _, l in OCN_TRACERS])
    sn, keep3 = _strs([n for n, _ in SED_TRACERS])
    sl, keep4 = _strs([l for _, l in SED_TRACERS])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    _ck(L.cg_restart_biogem_write(path.encode(), I, J, K, _ip(k1), *[_dp(a) for a in ax], len(OCN_TRACERS), on, ol, _dp(ocn),
                                  len(SED_TRACERS), sn, sl, _dp(part), float(year), run_id.encode()))
    return path


def read_biogem_restart(e, path, member=0, force_goldstein_ts=True, saln0=34.9):
    """sub_data_load_rst: ocn and bio_part of one member from a BIOGEM netCDF restart (any cGENIE run with the same grid; tracers
    the file does not hold keep their values), then the biogeochemical tracers of ts are rebuilt from ocn as initialise_biogem
    does (biogem.f90:495-508).  Returns the names found."""
    L = _lib.load()
    I, J, K = e.maxi, e.maxj, e.maxk
    k1 = np.ascontiguousarray(e.iconst("k1"), dtype=np.int32)
    ocn = np.ascontiguousarray(e.get("ocn", member), dtype=np.float64)
    part = np.ascontiguousarray(e.get("bio_part", member), dtype=np.float64)
    on, keep1 = _strs([n for n, _ in OCN_TRACERS])
    sn, keep3 = _strs([n for n, _ in SED_TRACERS])
    fo, fs = np.zeros(len(OCN_TRACERS), dtype=np.int32), np.zeros(len(SED_TRACERS), dtype=np.int32)
    _ck(L.cg_restart_biogem_read(path.encode(), I, J, K, _ip(k1), len(OCN_TRACERS), on, _dp(ocn), _ip(fo),
                                 len(SED_TRACERS), sn, _dp(part), _ip(fs)))
    e.put("ocn", ocn, member)
    e.put("bio_part", part, member)
    # sub_biogem_copy_ocntots (biogem_box.f90:3691-3739, ctrl_misc_Snorm): the biogeochemical tracers of GOLDSTEIN's ts
    # (and ts1) are the salinity-normalised ocn
    Lt = e.maxl
    o4 = ocn.reshape(K, J, I, Lt)
    k1ij = k1.reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet = np.arange(1, K + 1)[:, None, None] >= k1ij[None]
    V = np.asarray(e.const("bg_V"), dtype=np.float64).reshape(K, J, I)
    tot_V = np.cumsum(np.where(wet, V, 0.0).ravel())[-1]                      # sequential, in array order
    mean_S = np.cumsum(np.where(wet, o4[..., 1] * V, 0.0).ravel())[-1] / tot_V
    ts = np.ascontiguousarray(e.get("ts", member), dtype=np.float64).reshape(K, J, I, Lt)
    S = np.where(wet, o4[..., 1], 1.0)
    for l in range(2, Lt):
        ts[..., l] = np.where(wet, o4[..., l] * (mean_S / S), ts[..., l])
    if force_goldstein_ts:   # ctrl_force_GOLDSTEInTS (default .TRUE.): sub_biogem_copy_ocntotsTS, biogem_box.f90:3769-3790
        ts[..., 0] = np.where(wet, o4[..., 0] - 273.15, ts[..., 0])
        ts[..., 1] = np.where(wet, o4[..., 1] - saln0, ts[..., 1])
    e.put("ts", ts.ravel(), member)
    return [n for (n, _), f in zip(OCN_TRACERS, fo) if f] + [n for (n, _), f in zip(SED_TRACERS, fs) if f]
