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
    # inm_netcdf: u1(1:2) = the file's velocities, u = u1 (goldstein_data.f90:91-93); velc relaxes the new velocities
    # against u1 (goldstein.f90:3648-3654), and initialise_goldstein recomputes rho from the restored T, S
    e.put("u1", np.ascontiguousarray(u.reshape(K, J, I, 3)[..., :2]).ravel(), member)
    e.refresh_rho(member)
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


def _apply_biogem_restart(e, member, ocn, part, k1, force_goldstein_ts, saln0):
    """Stores a loaded ocn / bio_part and rebuilds GOLDSTEIN's ts from ocn as initialise_biogem does (biogem.f90:495-508)."""
    I, J, K = e.maxi, e.maxj, e.maxk
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
    if force_goldstein_ts:   # T, S changed under GOLDSTEIN: its density follows (the momentum step reads rho first)
        e.refresh_rho(member)


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
    _apply_biogem_restart(e, member, ocn, part, k1, force_goldstein_ts, saln0)
    return [n for (n, _), f in zip(OCN_TRACERS, fo) if f] + [n for (n, _), f in zip(SED_TRACERS, fs) if f]


# ------------------------------------------------------------------ ATCHEM (frozen selection: ia = 1,2,3,4,5,6,18,19)
# (ia, string_atm, string_longname_atm) from data/main/tracer_define.atm columns 2, 1 and 5
ATM_TRACERS = [(1, "temp", "surface air temperature"), (2, "humidity", "specific humidity"), (3, "pCO2", "carbon dioxide (CO2)"),
               (4, "pCO2_13C", "d13C CO2"), (5, "pCO2_14C", "d14C CO2"), (6, "pO2", "oxygen (O2)"), (18, "pCFC11", "CFC-11"),
               (19, "pCFC12", "CFC-12")]
# io / is of the frozen ocean and particulate selections (tracer_define.ocn / .sed column 2), in OCN_TRACERS / SED_TRACERS order
OCN_IDS = [1, 2, 3, 4, 5, 8, 10, 12, 15, 16, 17, 20, 35, 45, 46, 50]
SED_IDS = [3, 4, 5, 8, 14, 15, 16, 33, 34]


def atchem_axes(n_i, n_j):
    """phys_atm's lon / lat and their edges through edge_maker (atchem_data.f90:195-229, atchem_data_netCDF.f90:87-92): the
    sine-latitude grid is rebuilt from -pi/2 .. pi/2 there, not taken from GOLDSTEIN."""
    off = -260.0
    lon = np.array([(360.0 / n_i) * (float(i) - 0.5) + off for i in range(1, n_i + 1)])
    lone = np.array([(360.0 / n_i) * float(i) + off for i in range(1, n_i + 1)])
    lon_e = np.concatenate([[lone[0] - 360.0 / n_i], lone])
    s0, s1 = np.sin(-np.pi / 2), np.sin(np.pi / 2)
    ds = (s1 - s0) / float(n_j)
    sv = np.array([s0 + float(j) * ds for j in range(0, n_j + 1)])
    s = sv - 0.5 * ds
    rad = 180.0 / np.pi
    lat = rad * np.arcsin(s[1:])
    latn = rad * np.arcsin(sv[1:])
    lat_e = np.concatenate([[latn[0] - rad * (np.arcsin(sv[1]) - np.arcsin(sv[0]))], latn])
    return lon, lat, lon_e, lat_e


def _atm_check(e):
    if e.field_size("atm") != len(ATM_TRACERS) * e.maxi * e.maxj:
        raise RestartError("ATCHEM restart: the job's atmosphere tracer selection is not the frozen one")


def write_atchem_restart(e, path, member=0, year=0.0, run_id="", binary=False):
    """atchem_save_rst (atchem.f90:160-199): the netCDF restart of one member's atm array, or the unformatted dump."""
    L = _lib.load()
    _atm_check(e)
    I, J = e.maxi, e.maxj
    atm = np.ascontiguousarray(e.get("atm", member), dtype=np.float64)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if binary:
        ids = np.array([ia for ia, _, _ in ATM_TRACERS], dtype=np.int32)
        _ck(L.cg_restart_atchem_write_bin(path.encode(), I, J, len(ATM_TRACERS), _ip(ids), _dp(atm)))
        return path
    ax = [np.ascontiguousarray(a, dtype=np.float64) for a in atchem_axes(I, J)]
    an, keep1 = _strs([n for _, n, _ in ATM_TRACERS])
    al, keep2 = _strs([l for _, _, l in ATM_TRACERS])
    _ck(L.cg_restart_atchem_write(path.encode(), I, J, *[_dp(a) for a in ax], len(ATM_TRACERS), an, al, _dp(atm), float(year),
                                  run_id.encode()))
    return path


def read_atchem_restart(e, path, member=0, binary=False):
    """sub_data_load_rst (atchem_data.f90:89-189): the member's atm array from an ATCHEM restart; initialise_atchem hands the
    same values to the coupler as sfcatm (atchem.f90:54), and cpl_comp_atmocn copies tracers 3..n_atm of it to the ocean-grid
    sfcatm1 (atchem.f90:252-264), so those rows of sfcatm1 are rewritten too.  Returns the names found."""
    L = _lib.load()
    _atm_check(e)
    I, J = e.maxi, e.maxj
    atm = np.ascontiguousarray(e.get("atm", member), dtype=np.float64)
    fa = np.zeros(len(ATM_TRACERS), dtype=np.int32)
    if binary:
        ids = np.array([ia for ia, _, _ in ATM_TRACERS], dtype=np.int32)
        _ck(L.cg_restart_atchem_read_bin(path.encode(), I, J, len(ATM_TRACERS), _ip(ids), _dp(atm), _ip(fa)))
    else:
        an, keep1 = _strs([n for _, n, _ in ATM_TRACERS])
        _ck(L.cg_restart_atchem_read(path.encode(), I, J, len(ATM_TRACERS), an, _dp(atm), _ip(fa)))
    e.put("atm", atm, member)
    n = len(ATM_TRACERS)
    sfc = np.ascontiguousarray(e.get("sfcatm1", member), dtype=np.float64).reshape(J * I, n)
    sfc[:, 2:] = atm.reshape(J * I, n)[:, 2:]
    e.put("sfcatm1", sfc.ravel(), member)
    return [n for (_, n, _), f in zip(ATM_TRACERS, fa) if f]


def write_biogem_restart_bin(e, path, member=0):
    """biogem_save_restart, binary branch (biogem.f90:2340-2358): ocn and bio_part in full double precision."""
    L = _lib.load()
    if e.maxl != len(OCN_TRACERS):
        raise RestartError("BIOGEM restart: the job's tracer selection is not the frozen one")
    ocn = np.ascontiguousarray(e.get("ocn", member), dtype=np.float64)
    part = np.ascontiguousarray(e.get("bio_part", member), dtype=np.float64)
    io, isd = np.array(OCN_IDS, dtype=np.int32), np.array(SED_IDS, dtype=np.int32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    _ck(L.cg_restart_biogem_write_bin(path.encode(), e.maxi, e.maxj, e.maxk, len(io), _ip(io), _dp(ocn), len(isd), _ip(isd), _dp(part)))
    return path


def read_biogem_restart_bin(e, path, member=0, force_goldstein_ts=True, saln0=34.9):
    """sub_data_load_rst, binary branch (biogem_data.f90:540-550), then the same ts rebuild as read_biogem_restart."""
    L = _lib.load()
    k1 = np.ascontiguousarray(e.iconst("k1"), dtype=np.int32)
    ocn = np.ascontiguousarray(e.get("ocn", member), dtype=np.float64)
    part = np.ascontiguousarray(e.get("bio_part", member), dtype=np.float64)
    io, isd = np.array(OCN_IDS, dtype=np.int32), np.array(SED_IDS, dtype=np.int32)
    fo, fs = np.zeros(len(io), dtype=np.int32), np.zeros(len(isd), dtype=np.int32)
    _ck(L.cg_restart_biogem_read_bin(path.encode(), e.maxi, e.maxj, e.maxk, len(io), _ip(io), _dp(ocn), _ip(fo), len(isd), _ip(isd),
                                     _dp(part), _ip(fs)))
    _apply_biogem_restart(e, member, ocn, part, k1, force_goldstein_ts, saln0)
    return [n for (n, _), f in zip(OCN_TRACERS, fo) if f] + [n for (n, _), f in zip(SED_TRACERS, fs) if f]
