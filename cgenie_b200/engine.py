"""Python host mirror of the reference's module entry points for the hot path.

Method names, argument meaning and error behaviour follow the Fortran module procedures the coupler
calls through src/wrappers/genie_loop_wrappers.f90 (surflux, step_embm, step_seaice, step_goldstein,
...).  Everything forwards to the C-ABI of include/cgenie_b200.h; there is no Python or CPU compute
path.  A non-zero status raises CgenieError, the analogue of die()/write_status('ERRORED')
(src/wrappers/genie_util.f90:15-33).
"""
import ctypes as C

import numpy as np

from . import _lib


class CgenieError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cgenie_b200 error %d: %s" % (code, msg))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(_lib.D) if a is not None else None


class _Base:
    h = None

    def _ck(self, rc):
        if rc != 0:
            raise CgenieError(rc, self.L.cg_last_error().decode())

    def close(self):
        if self.h is not None:
            self.L.cg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state movement
    def field_size(self, name):
        n = self.L.cg_field_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        return n

    def get(self, name, member=0):
        """One member of a named field, flat in the reference's Fortran order."""
        out = np.empty(self.field_size(name), dtype=np.float64)
        self._ck(self.L.cg_sync_to_host(self.h, name.encode(), member, _dp(out), out.size))
        return out

    def put(self, name, values, member=0):
        a = np.ascontiguousarray(values, dtype=np.float64).ravel()
        self._ck(self.L.cg_sync_from_host(self.h, name.encode(), member, _dp(a), a.size))

    def get_all(self, name, out=None):
        """All members in the device-native layout [...][member_stride]."""
        n = self.field_size(name) * self.member_stride
        if out is None:
            out = np.empty(n, dtype=np.float64)
        self._ck(self.L.cg_sync_all_to_host(self.h, name.encode(), _dp(out), n))
        return out

    def put_all(self, name, values):
        a = np.ascontiguousarray(values, dtype=np.float64).ravel()
        self._ck(self.L.cg_sync_all_from_host(self.h, name.encode(), _dp(a), a.size))

    def wet_size(self, name):
        """doubles per member of a 3-D ocean field in the wet-cell packed exchange layout"""
        n = self.L.cg_wet_size(self.h, name.encode())
        if n < 0:
            raise CgenieError(1, self.L.cg_last_error().decode())
        return n

    def get_all_wet(self, name, out=None):
        """All members of a 3-D ocean field, wet cells only: [wet cell][inner][member_stride] (cg_sync_all_wet_to_host)."""
        n = self.wet_size(name) * self.member_stride
        if out is None:
            out = np.empty(n, dtype=np.float64)
        self._ck(self.L.cg_sync_all_wet_to_host(self.h, name.encode(), _dp(out), n))
        return out

    def put_all_wet(self, name, values):
        a = np.ascontiguousarray(values, dtype=np.float64).ravel()
        self._ck(self.L.cg_sync_all_wet_from_host(self.h, name.encode(), _dp(a), a.size))

    # ---- double-buffered exchange (cg_exchange_*): copies cross PCIe on a copy stream while the model computes
    def exchange_begin_upload(self, name, values, wet=False):
        """Start the asynchronous upload of all members of `name` (device layout; wet: wet cells only) from a page-locked array
        that stays untouched until exchange_wait()."""
        a = values if (isinstance(values, np.ndarray) and values.dtype == np.float64 and values.flags.c_contiguous) else np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.L.cg_exchange_begin_upload(self.h, name.encode(), 1 if wet else 0, a.ctypes.data, a.size))

    def exchange_commit_upload(self, name, also=None):
        """The staged upload of `name` becomes the field (and, with also=, a second field that takes the same data)."""
        self._ck(self.L.cg_exchange_commit_upload(self.h, name.encode(), also.encode() if also else None))

    def exchange_begin_download(self, name, out, wet=False):
        """Start the asynchronous download of all members of `name` into the page-locked array `out` (valid after exchange_wait())."""
        self._ck(self.L.cg_exchange_begin_download(self.h, name.encode(), 1 if wet else 0, out.ctypes.data, out.size))

    def exchange_wait(self):
        self._ck(self.L.cg_exchange_wait(self.h))

    def const(self, name):
        n = self.L.cg_const_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self._ck(self.L.cg_get_const(self.h, name.encode(), 0, _dp(out), n))
        return out

    def iconst(self, name):
        n = self.L.cg_const_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int32)
        self._ck(self.L.cg_get_iconst(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return out

    def _dims(self):
        d = (C.c_int32 * 8)()
        self._ck(self.L.cg_get_dims(self.h, d))
        (self.maxi, self.maxj, self.maxk, self.maxl, self.n_members, self.member_stride, self.nyear, self.ndta) = list(d)

    # ---- measurement
    def synchronize(self):
        self._ck(self.L.cg_synchronize(self.h))

    def launch_count(self, reset=False):
        return int(self.L.cg_launch_count(self.h, 1 if reset else 0))

    def timer_start(self):
        self._ck(self.L.cg_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_double()
        self._ck(self.L.cg_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def profile(self, on=True):
        self._ck(self.L.cg_profile_enable(self.h, 1 if on else 0))

    def profile_get(self, family):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.L.cg_profile_get(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_tracer_variant(self, variant):
        """'strict' = reference operation order (bit-exact vs the oracle); 'fast' = FMA + hoisted coefficients;
        'col' = fused flux + convection column kernel (same tolerance as 'fast')."""
        self._ck(self.L.cg_set_tracer_variant(self.h, {"strict": 0, "fast": 1, "col": 2}.get(variant, variant)))

    def tracer_variant_active(self):
        """'strict' / 'fast' / 'col': what the tracer step really runs ('col' exists for compiled grid shapes only)."""
        return ("strict", "fast", "col")[self.L.cg_tracer_variant_active(self.h)]

    def set_biogem_fusion(self, on):
        """run(): fuse biogem_tracercoupling's per-cell update into the step_biogem kernel (default off: slower)."""
        self._ck(self.L.cg_set_biogem_fusion(self.h, 1 if on else 0))

    def set_graphs(self, on):
        self._ck(self.L.cg_set_graphs(self.h, 1 if on else 0))

    def global_means(self):
        out = np.empty(self.n_members * self.maxl, dtype=np.float64)
        self._ck(self.L.cg_global_means(self.h, _dp(out)))
        return out.reshape(self.n_members, self.maxl)

    def health(self):
        out = np.empty(self.n_members, dtype=np.int32)
        self._ck(self.L.cg_health(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out


class Ensemble(_Base):
    """A batch of independent, parameter-perturbed cGENIE members resident on one GPU.

    jobdir     : job directory as written by the reference's new-job (namelists + input/<module>/ data)
    n_members  : ensemble size on this GPU
    perturb    : {param_name: array(n_members)} per-member overrides (cg_set_member_param)
    """

    def __init__(self, jobdir, n_members=1, device=0, perturb=None):
        self.L = _lib.load()
        h = _lib.P()
        self._ck(self.L.cg_create(str(jobdir).encode(), int(n_members), int(device), C.byref(h)))
        self.h = h
        for name, vals in (perturb or {}).items():
            a = np.ascontiguousarray(vals, dtype=np.float64)
            if a.size != n_members:
                raise ValueError("perturbation %s needs %d values" % (name, n_members))
            self._ck(self.L.cg_set_member_param(self.h, name.encode(), _dp(a)))
        self._ck(self.L.cg_initialise(self.h))
        self._dims()
        self.istep_ocn = self.istep_atm = self.istep_sic = 0

    # ---- module entry points (argument-less wrappers of genie_loop_wrappers.f90) ----
    def surflux(self, io=None):
        """surflux_wrapper (genie_loop_wrappers.f90:7-59): istep_ocn is incremented first (genie.f90:275)."""
        self.istep_ocn += 1
        self._ck(self.L.cg_surflux_step(self.h, self.istep_ocn, io))

    def step_embm(self, io=None):
        """embm_wrapper (:61-86)."""
        self.istep_atm += 1
        self._ck(self.L.cg_embm_step(self.h, self.istep_atm, io))

    def step_seaice(self, io=None):
        """gold_seaice_wrapper (:94-113)."""
        self.istep_sic += 1
        self._ck(self.L.cg_seaice_step(self.h, self.istep_sic, io))

    def step_goldstein(self, io=None):
        """goldstein_wrapper (:122-151)."""
        self._ck(self.L.cg_goldstein_step(self.h, self.istep_ocn, io))

    def biogem_init_ocn(self):
        """Build BIOGEM's ocn from the current ts (initialise_biogem, biogem.f90:283-285)."""
        self._ck(self.L.cg_biogem_init_ocn(self.h))

    def biogem_tracercoupling(self, go_ts=None, go_ts1=None):
        """biogem_tracercoupling_wrapper (genie_loop_wrappers.f90:318-322)."""
        self._ck(self.L.cg_biogem_tracercoupling(self.h, _dp(go_ts) if go_ts is not None else None,
                                                  _dp(go_ts1) if go_ts1 is not None else None))

    def biogem_climate(self):
        """biogem_climate_wrapper (genie_loop_wrappers.f90:324-345); resets the convection counter."""
        self._ck(self.L.cg_biogem_climate(self.h))

    def biogem_climate_sol(self):
        """biogem_climate_sol_wrapper (genie_loop_wrappers.f90:338-342): the insolation of the last surflux call; genie.f90
        calls it once, ahead of the very first BIOGEM step (genie.f90:369-370)."""
        self._ck(self.L.cg_biogem_climate_sol(self.h))

    def biogem_forcing(self, genie_clock_ms):
        """biogem_forcing_wrapper (genie_loop_wrappers.f90:324-328)."""
        self._ck(self.L.cg_biogem_forcing(self.h, int(genie_clock_ms)))

    def biogem_step(self, dts, genie_clock_ms):
        """biogem_wrapper (genie_loop_wrappers.f90:310-316); interface arrays stay on the device."""
        self._ck(self.L.cg_biogem_step(self.h, float(dts), int(genie_clock_ms)))

    def atchem_step(self, dts):
        """atchem_wrapper + cpl_comp_atmocn_wrapper (genie_loop_wrappers.f90:452-462)."""
        self._ck(self.L.cg_atchem_step(self.h, float(dts)))

    def biogem_sig_update(self, dts, ben_Dmin=0.0):
        """The ocean / atmosphere integrals of diag_biogem_timeseries (biogem.f90:2836-2917) for one BIOGEM step, on the device;
        call it on the steps of a save window, read the window with get("bg_sig", member)."""
        self._ck(self.L.cg_biogem_sig_update(self.h, float(dts), float(ben_Dmin)))

    def goldstein_mldta(self, member=0):
        """step_goldstein's go_mldta (goldstein.f90:449): mixed-layer depth in metres below the surface, (maxi,maxj); zero unless
        imld = 1."""
        out = np.zeros(self.maxi * self.maxj, dtype=np.float64)
        self._ck(self.L.cg_goldstein_mldta(self.h, int(member), _dp(out)))
        return out

    def biogem_sig_extended(self):
        """From now on step_biogem keeps sfxatm1 and the export through the surface layer's base, and biogem_sig_update also
        accumulates the sea-ice, overturning, land-temperature, export and air-sea flux integrals (field "bg_sig2";
        biogem.f90:2870-2883, 2926-2964, 3058-3062).  Call before the first BIOGEM step of interest."""
        self._ck(self.L.cg_biogem_sig_extended(self.h))

    def biogem_slice_update(self, dts):
        """diag_biogem_timeslice's arithmetic for one BIOGEM step of a save window: 3-D carbonate re-solve + window integrals
        (fields sl_ocn, sl_part, sl_carb, sl_carbconst, sl_carbisor, sl_t); call it behind biogem_climate (genie.f90:391-395)."""
        self._ck(self.L.cg_biogem_slice_update(self.h, float(dts)))

    def biogem_slice_reset(self):
        """sub_init_int_timeslice (biogem_data.f90:1012-1060)."""
        self._ck(self.L.cg_biogem_slice_reset(self.h))

    def biogem_sig_reset(self):
        """sub_init_int_timeseries (biogem_data.f90:964-1007)."""
        self._ck(self.L.cg_biogem_sig_reset(self.h))

    def cpl_flux_ocnsed(self, dts):
        """cpl_flux_ocnsed_wrapper (genie_loop_wrappers.f90:197-203): sfxsumsed += dts * sfxsed1 on the device."""
        self._ck(self.L.cg_cpl_flux_ocnsed(self.h, float(dts)))

    def cpl_comp_ocnsed(self, ocnstep, mbiogem, msedgem):
        """cpl_comp_ocnsed_wrapper (genie_loop_wrappers.f90:219-226): running mean of sfcocn1 into sfcsumocn."""
        self._ck(self.L.cg_cpl_comp_ocnsed(self.h, int(ocnstep), int(mbiogem), int(msedgem)))

    def reinit_flux_rokocn(self):
        """reinit_flux_rokocn_wrapper (genie_loop_wrappers.f90:289-293): sfxsumrok1 = 0."""
        self._ck(self.L.cg_reinit_flux_rokocn(self.h))

    def run(self, n_koverall):
        """n iterations of the genie.f90 main loop entirely on the device."""
        n = int(n_koverall)
        self._ck(self.L.cg_run(self.h, n))
        # keep the host-side step counters of the module-by-module interface in line (genie.f90:271-311)
        k0 = getattr(self, "koverall", 0)
        kocn = 5
        self.istep_atm += n
        self.istep_ocn += (k0 + n + kocn - 1) // kocn - (k0 + kocn - 1) // kocn     # iterations with MOD(k, kocn) == 1
        self.istep_sic += (k0 + n) // kocn - k0 // kocn                              # iterations with MOD(k, ksic) == 0
        self.koverall = k0 + n

    def set_koverall(self, koverall):
        """Restart support: continue the coupling loop from iteration `koverall` (multiple of kocn_loop)."""
        self._ck(self.L.cg_set_koverall(self.h, int(koverall)))
        kocn = 5
        self.istep_ocn = int(koverall) // kocn
        self.istep_sic = int(koverall) // kocn
        self.istep_atm = int(koverall)
        self.koverall = int(koverall)

    def refresh_rho(self, member=-1):
        """rho = eos(T, S) at the wet cells of one member (all members: -1) after T, S were rewritten from the host, as
        initialise_goldstein does behind inm_netcdf (goldstein.f90:1724-1760)."""
        self._ck(self.L.cg_refresh_rho(self.h, int(member)))

    def run_years(self, years):
        self.run(int(round(years * self.nyear * self.ndta)))


class TracerStep(_Base):
    """Stand-alone tstepo (flux + convection) on caller-provided fields (BASELINE config #5, kernel tests)."""

    def __init__(self, maxi, maxj, maxk, maxl, k1, n_members=1, device=0, diff1=2000.0, diff2=1.0e-5, nyear=96):
        self.L = _lib.load()
        k1 = np.ascontiguousarray(k1, dtype=np.int32).ravel()
        if k1.size != (maxi + 2) * (maxj + 2):
            raise ValueError("k1 must be (0:maxi+1, 0:maxj+1)")
        h = _lib.P()
        self._ck(self.L.cg_tracer_create(maxi, maxj, maxk, maxl, n_members, device, k1.ctypes.data_as(C.POINTER(C.c_int32)),
                                         diff1, diff2, nyear, C.byref(h)))
        self.h = h
        self._dims()

    def set(self, ts=None, u=None, tsflux=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (ts, u, tsflux)]
        self._ck(self.L.cg_tracer_set(self.h, _dp(a[0]), _dp(a[1]), _dp(a[2])))

    def step(self, n=1):
        self._ck(self.L.cg_tracer_step(self.h, int(n)))

    def fetch(self):
        M, I, J, K, Lt = self.n_members, self.maxi, self.maxj, self.maxk, self.maxl
        ts = np.empty(M * Lt * I * J * K)
        rho = np.empty(M * I * J * K)
        cost = np.empty(M * I * J)
        self._ck(self.L.cg_tracer_get(self.h, _dp(ts), _dp(rho), _dp(cost)))
        return ts.reshape(M, K, J, I, Lt), rho.reshape(M, K, J, I), cost.reshape(M, J, I)


class EnsembleGroups:
    """A shard of MORE than `group` members on one GPU, run as independent groups of `group` (<= 128) members: one library
    handle per group (own state, streams, captured graphs), all groups driven concurrently from host threads.

    The production tracer kernels are compiled for a member stride of up to 128 (one thread block = one wet column x 128
    members); members are independent, so a larger shard is simply several such ensembles side by side -- measured on
    B200: 2 x 128 members 6.8 M model-years/hour, 4 x 128 6.6 M, against 6.4 M for one group and 4.2 M for a single
    512-member handle (which falls back to the generic-shape kernels).  Member m lives in group m // group, lane
    m % group.  The interface is the subset of `Ensemble` that works on the whole shard."""

    def __init__(self, jobdir, n_members, device=0, perturb=None, group=128):
        import threading
        self._threading = threading
        if group < 1 or group > 128:
            raise ValueError("group size must be in 1..128")
        self.group = int(group)
        self.n_members = int(n_members)
        self.parts = []
        for g0 in range(0, self.n_members, self.group):
            n = min(self.group, self.n_members - g0)
            pert = {k: np.asarray(v, dtype=np.float64)[g0:g0 + n] for k, v in (perturb or {}).items()}
            self.parts.append(Ensemble(jobdir, n_members=n, device=device, perturb=pert))
        e0 = self.parts[0]
        (self.maxi, self.maxj, self.maxk, self.maxl, self.nyear, self.ndta) = (e0.maxi, e0.maxj, e0.maxk, e0.maxl, e0.nyear, e0.ndta)

    def close(self):
        for e in self.parts:
            e.close()
        self.parts = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _each(self, fn):
        """fn(part) for every group, concurrently (the C-ABI calls release the GIL); re-raises the first error."""
        if len(self.parts) == 1:
            return [fn(self.parts[0])]
        out, err = [None] * len(self.parts), []

        def work(q):
            try:
                out[q] = fn(self.parts[q])
            except Exception as ex:      # noqa: BLE001
                err.append(ex)
        th = [self._threading.Thread(target=work, args=(q,)) for q in range(len(self.parts))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise err[0]
        return out

    def set_tracer_variant(self, variant):
        for e in self.parts:
            e.set_tracer_variant(variant)

    def tracer_variant_active(self):
        return self.parts[0].tracer_variant_active()

    def run(self, n_koverall):
        self._each(lambda e: e.run(n_koverall))

    def synchronize(self):
        for e in self.parts:
            e.synchronize()

    def launch_count(self, reset=False):
        return sum(e.launch_count(reset) for e in self.parts)

    def timer_start(self):
        for e in self.parts:
            e.timer_start()

    def timer_stop_ms(self):
        """device time from the first group's start to the last group's end is not observable with per-handle events: the
        longest group (all were started together) is reported"""
        return max(e.timer_stop_ms() for e in self.parts)

    def health(self):
        return np.concatenate([e.health() for e in self.parts])

    def global_means(self):
        return np.concatenate([e.global_means() for e in self.parts], axis=0)

    def get(self, name, member=0):
        return self.parts[member // self.group].get(name, member % self.group)

    def put(self, name, values, member=0):
        self.parts[member // self.group].put(name, values, member % self.group)

    def field_size(self, name):
        return self.parts[0].field_size(name)

    def iconst(self, name):
        return self.parts[0].iconst(name)

    def const(self, name):
        return self.parts[0].const(name)
