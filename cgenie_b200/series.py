"""BIOGEM's ASCII time series (biogem_series_ocn_*.res, biogem_series_atm_*.res) of the frozen eb_go_gs_ac_bg tracer selection.

Python mirror of sub_init_data_save_runtime / sub_data_save_runtime (src/biogem/biogem_data_ascii.f90:23-110, 669-935) over the
C-ABI: the window integrals come from the device (Ensemble.biogem_sig_update, field "bg_sig"), the files are written by
cg_biogem_series_write (csrc/cg_series.cpp) in the reference's edit descriptors.  No arithmetic happens here."""
import ctypes as C
import os

import numpy as np

from . import _lib
from .restart import ATM_TRACERS, OCN_TRACERS, _strs

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
