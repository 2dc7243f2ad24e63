"""BIOGEM's ASCII time series (biogem_series_ocn_*.res, biogem_series_atm_*.res) of the frozen eb_go_gs_ac_bg tracer selection.

Python mirror of sub_init_data_save_runtime / sub_data_save_runtime (src/biogem/biogem_data_ascii.f90:23-110, 669-935) over the
C-ABI: the window integrals come from the device (Ensemble.biogem_sig_update, field "bg_sig"), the files are written by
cg_biogem_series_write (csrc/cg_series.cpp) in the reference's edit descriptors.  No arithmetic happens here."""
import ctypes as C
import os

import numpy as np

from . import _lib
from .restart import ATM_TRACERS, OCN_TRACERS, SED_TRACERS, _strs, biogem_axes

# ocn_type / ocn_dep (0-based compact index of the bulk tracer) and atm_type / atm_dep of the frozen selection
# (data/main/tracer_define.ocn, tracer_define.atm columns 4 and 3)
OCN_TYPE = [0, 0, 1, 11, 12, 1, 1, 1, 1, 11, 12, 1, 1, 1, 1, 1]
OCN_DEP = [0, 1, 2, 2, 2, 5, 6, 7, 8, 8, 8, 11, 12, 13, 14, 15]
ATM_TYPE = [0, 0, 1, 11, 12, 1, 1, 1]
ATM_DEP = [0, 1, 2, 2, 2, 5, 6, 7]


class SeriesError(RuntimeError):
    pass


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a.ctypes.data_as(C.POINTER(C.c_int32)), a


def write_series(outdir, sig=None, t_yr=0.0, outfile_name="biogem", with_sur=True):
    """sig=None: create the files with their header lines (sub_init_data_save_runtime); otherwise append the line of one save
    window from the integrals of one member (Ensemble.get("bg_sig", member))."""
    L = _lib.load()
    os.makedirs(outdir, exist_ok=True)
    on, k1 = _strs([n for n, _ in OCN_TRACERS])
    an, k2 = _strs([n for _, n, _ in ATM_TRACERS])
    ot, k3 = _i32(OCN_TYPE); od, k4 = _i32(OCN_DEP); at, k5 = _i32(ATM_TYPE); ad, k6 = _i32(ATM_DEP)
    sp = None
    if sig is not None:
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        if sig.size != 3 + 3 * len(OCN_TRACERS) + len(ATM_TRACERS):
            raise SeriesError("series: the integrals are not those of the frozen tracer selection")
        sp = sig.ctypes.data_as(_lib.D)
    rc = L.cg_biogem_series_write(str(outdir).encode(), outfile_name.encode(), 1 if sig is None else 0, float(t_yr), len(OCN_TRACERS), on,
                                  ot, od, len(ATM_TRACERS), an, at, ad, sp, 1 if with_sur else 0)
    if rc:
        raise SeriesError(L.cg_series_last_error().decode())


# sed_type / sed_dep of the frozen particulate selection (tracer_define.sed columns 4 and 3; dep as 0-based compact index)
SED_TYPE = [1, 11, 12, 3, 1, 11, 12, 9, 9]
SED_DEP = [0, 0, 0, 3, 4, 4, 4, 7, 8]
ATLANTIC_TOPOS = ("worbe2", "worjh2", "worjh4", "worlg2", "worlg4", "wv2jh2", "wv3jh2", "worri4")     # biogem_data_ascii.f90:1282


def write_series_ext(outdir, sig=None, sig2=None, t_yr=0.0, outfile_name="biogem", ocn_tot_A=1.0, opsi_scale=1592.5, atlantic=True):
    """The fexport_*, fseaair_*, focnatm_*, misc_seaice, misc_opsi, misc_atm_D14C, misc_SLT files (cg_biogem_series_write_ext):
    sig = sig2 = None creates them with their header lines, otherwise one line per file from "bg_sig" and "bg_sig2"."""
    from .restart import SED_TRACERS
    L = _lib.load()
    os.makedirs(outdir, exist_ok=True)
    sn, k1 = _strs([n for n, _ in SED_TRACERS])
    an, k2 = _strs([n for _, n, _ in ATM_TRACERS])
    st, k3 = _i32(SED_TYPE); sd, k4 = _i32(SED_DEP); at, k5 = _i32(ATM_TYPE); ad, k6 = _i32(ATM_DEP)
    sp = sp2 = None
    if sig is not None:
        sig = np.ascontiguousarray(sig, dtype=np.float64)
        sig2 = np.ascontiguousarray(sig2, dtype=np.float64)
        if sig.size != 3 + 3 * len(OCN_TRACERS) + len(ATM_TRACERS) or sig2.size != 8 + len(SED_TRACERS) + 2 * len(ATM_TRACERS):
            raise SeriesError("series: the integrals are not those of the frozen tracer selection")
        sp, sp2 = sig.ctypes.data_as(_lib.D), sig2.ctypes.data_as(_lib.D)
    rc = L.cg_biogem_series_write_ext(str(outdir).encode(), outfile_name.encode(), 1 if sig is None else 0, float(t_yr), len(OCN_TRACERS),
                                      len(SED_TRACERS), sn, st, sd, len(ATM_TRACERS), an, at, ad, sp, sp2, float(ocn_tot_A),
                                      float(opsi_scale), 1 if atlantic else 0)
    if rc:
        raise SeriesError(L.cg_series_last_error().decode())


YR_S = 3600.0 * (24.0 * 365.25)          # conv_yr_s, gem_cmn.f90:511-513
S_YR = 1.0 / YR_S                        # conv_s_yr, gem_cmn.f90:532
NULLSMALL = 0.999999e-19                 # const_real_nullsmall, gem_cmn.f90:719
N_DATA_MAX = 32767                       # biogem_lib.f90:504

# ---------------------------------------------------------------------------------------------------------------- time slices
# units and valid range of the frozen ocean selection (data/main/tracer_define.ocn columns 6 - 8)
OCN_UNITS = ["degrees C", "PSU", "mol kg-1", "o/oo", "o/oo"] + ["mol kg-1"] * 4 + ["o/oo", "o/oo"] + ["mol kg-1"] * 5
OCN_MIMA = [(-9.999, 99.999), (0.0, 99.999), (-9.99E+2, 9.99E-1), (-9.99E+2, 9.99E+2), (-9.99E+5, 9.99E+5)] + \
    [(-9.99E+2, 9.99E-1)] * 4 + [(-9.99E+2, 9.99E+2), (-9.99E+5, 9.99E+5)] + [(-9.99E+2, 9.99E-1)] * 5
# rows of the device's "sl_carb" / "sl_carbconst" (include/cgenie_b200.h) and the reference's order of variables
# (string_carb / string_carbconst, gem_cmn.f90:425-455)
SL_CARB = ["H", "conc_CO2", "conc_CO3", "conc_HCO3", "fug_CO2", "ohm_cal", "ohm_arg", "dCO3_cal", "dCO3_arg", "RF0"]
SL_CARBCONST = ["k1", "k2", "k", "kB", "kW", "kSi", "kHF", "kHSO4", "kP1", "kP2", "kP3", "kH2S", "kNH4", "kcal", "karg", "QCO2", "QO2"]
REF_CARB = ["H", "fug_CO2", "conc_CO2", "conc_CO3", "conc_HCO3", "ohm_cal", "ohm_arg", "dCO3_cal", "dCO3_arg", "RF0"]
REF_CARBCONST = ["k", "k1", "k2", "kB", "kW", "kSi", "kHF", "kHSO4", "kP1", "kP2", "kP3", "kcal", "karg", "QCO2", "QO2", "kH2S", "kNH4"]


def ocean_mass(e):
    """phys_ocn(ipo_M) (biogem_data.f90:1126-1130): conv_m3_kg * dD * A at the wet cells, zero elsewhere; flat (n_i,n_j,n_k)."""
    I, J, K = e.maxi, e.maxj, e.maxk
    k1 = np.asarray(e.iconst("k1")).reshape(J + 2, I + 2)
    sv, dz = e.const("sv"), e.const("dz")
    m = np.zeros((K, J, I))
    for j in range(1, J + 1):
        A = 2.0 * 3.141592653589793 * (6.37e6 * 6.37e6) * (1.0 / I) * (sv[j] - sv[j - 1])
        for k in range(1, K + 1):
            m[k - 1, j - 1, :] = np.where(k1[j, 1:I + 1] <= k, 1027.649 * ((5.0e3 * dz[k]) * A), 0.0)
    return m.ravel()


def write_timeslice_3d(e, path, year_mid, member=0, run_id="", derived=True, carbconst=False):
    """One record of fields_biogem_3d.nc from the device's window integrals of one member (sub_save_netcdf + sub_save_netcdf_3d,
    biogem_data_netCDF.f90:282-459, 1959-2315; the caller resets the integrals afterwards as sub_init_int_timeslice does).
    derived = ctrl_data_save_derived (the _Snorm / _tot / bio_part_ blocks), carbconst = ctrl_data_save_slice_carbconst."""
    L = _lib.load()
    I, J, K = e.maxi, e.maxj, e.maxk
    if e.maxl != len(OCN_TRACERS):
        raise SeriesError("time slices: the job's tracer selection is not the frozen one")
    k1 = np.ascontiguousarray(e.iconst("k1"), dtype=np.int32)
    ax = [np.ascontiguousarray(a, dtype=np.float64) for a in biogem_axes(I, J, K, e.const("s"), e.const("sv"), e.const("dz"), e.const("dza"))]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    n3 = I * J * K
    ocn = np.ascontiguousarray(e.get("sl_ocn", member))
    part = np.ascontiguousarray(e.get("sl_part", member))
    carb = np.ascontiguousarray(e.get("sl_carb", member).reshape(n3, len(SL_CARB))[:, [SL_CARB.index(n) for n in REF_CARB]])
    cc = np.ascontiguousarray(e.get("sl_carbconst", member).reshape(n3, len(SL_CARBCONST))[:, [SL_CARBCONST.index(n) for n in REF_CARBCONST]])
    t = float(e.get("sl_t", member)[0])
    on, k_1 = _strs([n for n, _ in OCN_TRACERS]); ol, k_2 = _strs([l for _, l in OCN_TRACERS]); ou, k_3 = _strs(OCN_UNITS)
    mima = np.ascontiguousarray(OCN_MIMA, dtype=np.float64)
    ot, k_4 = _i32(OCN_TYPE); od, k_5 = _i32(OCN_DEP)
    sn, k_6 = _strs([n for n, _ in SED_TRACERS]); st, k_7 = _i32(SED_TYPE); sd, k_8 = _i32(SED_DEP)
    cn, k_9 = _strs(REF_CARB); ccn, k_10 = _strs(REF_CARBCONST)
    mass = np.ascontiguousarray(ocean_mass(e)) if derived else None
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    rc = L.cg_slice_biogem_write_3d(path.encode(), I, J, K, k1.ctypes.data_as(C.POINTER(C.c_int32)), *[dp(a) for a in ax],
                                    len(OCN_TRACERS), on, ol, ou, dp(mima), ot, od, dp(ocn), len(SED_TRACERS), sn, st, sd, dp(part),
                                    len(REF_CARB), cn, dp(carb), len(REF_CARBCONST) if carbconst else 0, ccn, dp(cc),
                                    dp(mass) if derived else None, t, float(year_mid), run_id.encode())
    if rc != 0:
        raise SeriesError(L.cg_restart_last_error().decode())
    return path

class SliceSaver:
    """The save-window logic of BIOGEM's time slices for ctrl_misc_t_BP = .FALSE.: the list of slice mid-points of sub_init_data_save
    (biogem_data.f90:2528-2584: the dates of biogem_save_timeslice.dat through sub_load_data_t1 as "years to go", the first one inside
    the run, ctrl_data_save_slice_autoend) and the window tests of diag_biogem_timeslice (biogem.f90:2451-2466, 2608-2696).  Scalar
    bookkeeping only: the integrals are the device's (Ensemble.biogem_slice_update), the file cg_slice_biogem_write_3d's.

    Call step(dts, genie_clock_ms) where genie.f90 calls diag_biogem_timeslice_wrapper (genie.f90:391-395): behind biogem_climate,
    ahead of the time-series update and atchem_step."""

    def __init__(self, e, path, t_runtime, save_times, t_start=0.0, slice_dt=1.0, slice_n=0, member=0, autoend=False, run_id="",
                 derived=True, carbconst=False):
        self.e, self.path, self.member, self.run_id, self.derived, self.carbconst = e, str(path), member, run_id, derived, carbconst
        self.t_runtime, self.t_end = float(t_runtime), float(t_start) + float(t_runtime)
        self.slice_dt, self.slice_n = float(slice_dt), int(slice_n)       # par_data_save_slice_dt, par_data_save_slice_n
        t_err = 3600.0 * 1.0 / YR_S
        data = [float(x) for x in (save_times or [])]
        n = len(data)
        if n and data[-1] <= data[0]:            # sub_load_data_t1, .NOT. ctrl_misc_t_BP (biogem_lib.f90:1487-1543)
            ts = [self.t_end - x for x in data]
        else:
            ts = [self.t_end - data[n - q] for q in range(1, n + 1)]
        i = n
        while i > 0:
            if ts[i - 1] < (self.t_runtime - self.slice_dt / 2.0 + t_err):
                break
            i -= 1
        if i > 0 and ts[i - 1] < (self.slice_dt / 2.0 - t_err):
            i = 0
        if autoend:
            if i == 0:
                ts, i = [self.slice_dt / 2.0], 1
            else:
                for q in range(i, 0, -1):
                    if ts[q - 1] < (1.0 - self.slice_dt / 2.0 + t_err):
                        if not ts[q - 1] > (self.slice_dt / 2.0 - t_err):
                            ts[q - 1] = self.slice_dt / 2.0
                        break
        self.ts, self.i = ts, i
        self.int_t, self.int_t_tot, self.count = 0.0, 0.0, 0
        self.saved = []
        e.biogem_slice_reset()

    def step(self, dts, genie_clock_ms):
        """One BIOGEM step.  Returns the year a record was written for, else None."""
        if self.i <= 0:
            return None
        s_yr = 1.0 / YR_S                                                    # conv_s_yr
        loc_t = self.t_runtime - float(genie_clock_ms) / (1000.0 * YR_S)
        if not (loc_t - (self.ts[self.i - 1] + self.slice_dt / 2.0)) < -s_yr:
            return None
        dtyr = float(dts) / YR_S
        self.e.biogem_slice_update(dts)
        self.int_t += dtyr
        self.int_t_tot += dtyr
        self.count += 1
        full = (self.slice_dt - self.int_t) < s_yr
        if not (full or self.count == self.slice_n):
            return None
        yr = self.t_end - loc_t - self.int_t / 2.0
        yr = float(int(yr)) + float(int(1000.0 * (yr - float(int(yr)) + 0.0005))) / 1000.0
        write_timeslice_3d(self.e, self.path, yr, member=self.member, run_id=self.run_id, derived=self.derived, carbconst=self.carbconst)
        self.saved.append(yr)
        self.e.biogem_slice_reset()                                          # sub_init_int_timeslice
        self.int_t, self.count = 0.0, 0
        if (self.slice_dt - self.int_t_tot) < s_yr:
            self.int_t_tot = 0.0
            self.i -= 1
        return yr


class SeriesSaver:
    """The save-window logic of BIOGEM's time series for ctrl_misc_t_BP = .FALSE.: sub_init_data_save (biogem_data.f90:2449-2527,
    the list of save times either from biogem_save_sig.dat through sub_load_data_t1, biogem_lib.f90:1487-1543, or generated
    from par_data_save_sig_dt) and the window tests of diag_biogem_timeseries (biogem.f90:2757-2769, 3079-3156).  Scalar
    bookkeeping only: the sums are the device's (Ensemble.biogem_sig_update), the files cg_biogem_series_write's.

    Call step(dts, genie_clock_ms) where genie.f90 calls diag_biogem_timeseries_wrapper (genie.f90:401-405): behind
    biogem_climate, ahead of atchem_step.  (Behind atchem_step the ocean rows are the same and the atmosphere rows one coupling
    interval newer; device and oracle agree at either point.)"""

    def __init__(self, e, outdir, t_runtime, t_start=0.0, sig_dt=1.0, save_times=None, ben_Dmin=0.0, member=0, with_sur=True,
                 autoend=False, outfile_name="biogem", extended=False, world="worjh2"):
        self.e, self.outdir, self.member, self.with_sur, self.outfile_name = e, str(outdir), member, with_sur, outfile_name
        self.t_runtime, self.t_end = float(t_runtime), float(t_start) + float(t_runtime)
        self.sig_dt, self.ben_Dmin = float(sig_dt), float(ben_Dmin)
        t_err = 3600.0 * 1.0 / YR_S      # par_misc_t_err, biogem_data.f90:393
        data = [float(x) for x in (save_times or [])]
        if data:                         # sub_load_data_t1, .NOT. ctrl_misc_t_BP: times become "years to go", ascending
            n = len(data)
            if data[-1] <= data[0]:
                sig = [self.t_end - x for x in data]
            else:
                sig = [self.t_end - data[n - q] for q in range(1, n + 1)]
        else:
            if not self.sig_dt > NULLSMALL:
                raise SeriesError("time-series save interval must be non-zero and positive")
            n = int(self.t_runtime / self.sig_dt + NULLSMALL)
            while n > N_DATA_MAX:
                self.sig_dt = 10.0 * self.sig_dt
                n = int(self.t_runtime / self.sig_dt + NULLSMALL)
            sig = [float(q - 0.5) * self.sig_dt + (self.t_runtime - float(n) * self.sig_dt) for q in range(1, n + 1)]
        i = len(sig)
        while i > 0:                     # the first save point that lies inside the run
            if sig[i - 1] < (self.t_runtime - self.sig_dt / 2.0 + t_err):
                break
            i -= 1
        if autoend:                      # ctrl_data_save_sig_autoend
            for q in range(i, 0, -1):
                if sig[q - 1] < (1.0 - self.sig_dt / 2.0 + t_err):
                    if not sig[q - 1] > (self.sig_dt / 2.0 - t_err):
                        sig[q - 1] = self.sig_dt / 2.0
                    break
        self.sig, self.sig_i = sig, i
        self.int_t_sig = 0.0
        self.saved = []
        write_series(self.outdir, None, outfile_name=outfile_name, with_sur=with_sur)     # sub_init_data_save_runtime
        self.extended = bool(extended)
        if self.extended:                # the export, air-sea flux and "misc" series as well
            e.biogem_sig_extended()
            self.ext_kw = dict(outfile_name=outfile_name, ocn_tot_A=float(e.const("bg_ocn_tot_A")[0]), atlantic=world in ATLANTIC_TOPOS)
            write_series_ext(self.outdir, **self.ext_kw)
        e.biogem_sig_reset()

    def step(self, dts, genie_clock_ms):
        loc_t = self.t_runtime - float(genie_clock_ms) / (1000.0 * YR_S)
        dtyr = float(dts) / YR_S
        if not self.sig_i > 0:
            return
        if (loc_t - (self.sig[self.sig_i - 1] + self.sig_dt / 2.0)) < -S_YR:      # inside the window
            self.e.biogem_sig_update(dts, self.ben_Dmin)
            self.int_t_sig = self.int_t_sig + dtyr          # the device adds the same dtyr to its int_t_sig
        if (self.sig_dt - self.int_t_sig) < S_YR:           # the window is full: save and move to the next save point
            yr = self.t_end - loc_t - self.int_t_sig / 2.0
            yr = float(int(yr)) + float(int(1000.0 * (yr - float(int(yr)) + 0.0005))) / 1000.00
            if self.int_t_sig > NULLSMALL:
                write_series(self.outdir, self.e.get("bg_sig", self.member), t_yr=yr, outfile_name=self.outfile_name,
                             with_sur=self.with_sur)
                if self.extended:
                    write_series_ext(self.outdir, self.e.get("bg_sig", self.member), self.e.get("bg_sig2", self.member), t_yr=yr,
                                     **self.ext_kw)
                self.saved.append(yr)
            self.sig_i -= 1
            self.e.biogem_sig_reset()                       # sub_init_int_timeseries
            self.int_t_sig = 0.0
