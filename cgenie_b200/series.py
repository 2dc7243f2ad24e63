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
