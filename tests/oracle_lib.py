"""ctypes binding of the CPU oracle (oracle/libcgo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
product package cgenie_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "libcgo_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        P = C.c_void_p
        L.cgo_create.restype = P
        L.cgo_create.argtypes = [C.c_char_p] + [P] * 2 + [C.c_int] + [P] * 8
        L.cgo_destroy.argtypes = [P]
        for f in ("cgo_surflux", "cgo_embm_step", "cgo_seaice_step", "cgo_goldstein_step", "cgo_tstepo",
                  "cgo_tstepo_flux", "cgo_co", "cgo_momentum", "cgo_tstipa", "cgo_biogem_init",
                  "cgo_biogem_tracercoupling"):
            if hasattr(L, f):
                getattr(L, f).argtypes = [P]
                getattr(L, f).restype = None
        for f in ("cgo_biogem_forcing", "cgo_biogem_climate", "cgo_cpl_flux_ocnatm", "cgo_atchem_step"):
            getattr(L, f).argtypes = [P]
            getattr(L, f).restype = None
        L.cgo_cpl_flux_ocnsed.argtypes, L.cgo_cpl_flux_ocnsed.restype = [P, C.c_double], None
        L.cgo_cpl_comp_ocnsed.argtypes, L.cgo_cpl_comp_ocnsed.restype = [P, C.c_int, C.c_int, C.c_int], None
        L.cgo_reinit_flux_rokocn.argtypes, L.cgo_reinit_flux_rokocn.restype = [P], None
        L.cgo_biogem_sig_update.argtypes, L.cgo_biogem_sig_update.restype = [P, C.c_double], None
        L.cgo_biogem_sig_auto.argtypes, L.cgo_biogem_sig_auto.restype = [P, C.c_int, C.c_double], None
        L.cgo_biogem_slice_update.argtypes, L.cgo_biogem_slice_update.restype = [P], None
        L.cgo_biogem_slice_auto.argtypes, L.cgo_biogem_slice_auto.restype = [P, C.c_int], None
        L.cgo_biogem_step.argtypes = [P]
        L.cgo_biogem_step.restype = C.c_int
        L.cgo_biogem_setup.argtypes = [P, C.c_char_p]
        L.cgo_biogem_setup.restype = None
        L.cgo_run.argtypes = [P, C.c_long]
        L.cgo_field.restype = C.POINTER(C.c_double)
        L.cgo_field.argtypes = [P, C.c_char_p, C.POINTER(C.c_long)]
        L.cgo_ifield.restype = C.POINTER(C.c_int)
        L.cgo_ifield.argtypes = [P, C.c_char_p, C.POINTER(C.c_long)]
        L.cgo_scalar.restype = C.c_double
        L.cgo_scalar.argtypes = [P, C.c_char_p]
        L.cgo_set_scalar.argtypes = [P, C.c_char_p, C.c_double]
        _LIB = L
    return _LIB


def load_inputs(world):
    z = np.load(os.path.join(ROOT, "configs", "inputs.npz"))
    d = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(world + "/")}
    d.update({k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("winds/")})
    return d


class Oracle:
    """One ensemble member of the CPU restatement.  world=None + k1=(maxj+2, maxi+2) file-order array gives a
    stand-alone tracer-step oracle on a synthetic grid (no atmosphere, no barotropic set-up)."""

    def __init__(self, world="worbe2", k1=None, **params):
        self.L = lib()
        keep = []
        if world is None:
            params = dict(params, tracer_only=1)
            kv = "".join("%s=%r\n" % (k, float(v)) for k, v in params.items())
            k1 = np.ascontiguousarray(k1, dtype=np.int32)
            I, J = int(params["maxi"]), int(params["maxj"])
            ps = np.zeros((J + 1) * I)
            npi = np.zeros(1, dtype=np.int32)
            keep += [k1, ps, npi]
            self.h = self.L.cgo_create(kv.encode(), k1.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p), 0,
                                       npi.ctypes.data_as(C.c_void_p), None, None, None, None, None, None, None)
            if not self.h:
                raise RuntimeError("cgo_create failed")
            self.params = params
            self._keep = keep
            return
        inp = load_inputs(world)
        kv = "".join("%s=%r\n" % (k, float(v)) for k, v in params.items())

        def ptr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        self.h = self.L.cgo_create(kv.encode(), ptr(inp["k1"], np.int32), ptr(inp["psiles"], np.float64),
                                   len(inp["npi"]), ptr(inp["npi"], np.int32), ptr(inp["paths"], np.int32),
                                   ptr(inp["taux_u"], np.float64), ptr(inp["tauy_u"], np.float64),
                                   ptr(inp["taux_v"], np.float64), ptr(inp["tauy_v"], np.float64),
                                   ptr(inp["uncep"], np.float64), ptr(inp["vncep"], np.float64))
        if not self.h:
            raise RuntimeError("cgo_create failed")
        self.params = params

    def biogem_setup(self, **bg_params):
        """initialise_biogem + initialise_atchem for the frozen eb_go_gs_ac_bg configuration (oracle/cgo_biogem.c);
        afterwards run() also executes the BIOGEM/ATCHEM block of genie.f90.  bg_params override data_BIOGEM keys."""
        z = np.load(os.path.join(ROOT, "configs", "inputs.npz"))
        self.f("bg_windspeed")[:] = z["biogem/worjh2_windspeed"][::-1, :].ravel()   # file rows j = maxj..1
        kv = "".join("%s=%r\n" % (k, float(v)) for k, v in dict(self.params, **bg_params).items())
        self.L.cgo_biogem_setup(self.h, kv.encode())

    def close(self):
        if self.h:
            self.L.cgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def f(self, name):
        """numpy view (no copy) of a double field, flat Fortran order."""
        n = C.c_long()
        p = self.L.cgo_field(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def i(self, name):
        n = C.c_long()
        p = self.L.cgo_ifield(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def s(self, name):
        return self.L.cgo_scalar(self.h, name.encode())

    def set(self, name, v):
        self.L.cgo_set_scalar(self.h, name.encode(), float(v))

    def run(self, nkoverall):
        self.L.cgo_run(self.h, int(nkoverall))

    def call(self, fn):
        getattr(self.L, "cgo_" + fn)(self.h)
