"""ctypes loader for libcgenie_b200.so (the C-ABI of include/cgenie_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgenie_b200.so")
_LIB = None

P = C.c_void_p
D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)
STRS = C.POINTER(C.c_char_p)


class SurfluxIO(C.Structure):
    _fields_ = [(n, D) for n in (
        "albedo_ocn", "latent_ocn", "sensible_ocn", "netsolar_ocn", "netlong_ocn", "evap_ocn", "precip_ocn",
        "runoff_ocn", "runoff_land", "latent_atm", "sensible_atm", "netsolar_atm", "netlong_atm", "evap_atm",
        "precip_atm", "dhght_sic", "dfrac_sic", "temp_sic", "albd_sic", "qstar_atm")]


class EmbmIO(C.Structure):
    _fields_ = [(n, D) for n in ("tstar_atm", "qstar_atm")]


class SeaiceIO(C.Structure):
    _fields_ = [(n, D) for n in ("hght_sic", "frac_sic", "waterflux_ocn", "conductflux_ocn")]


class GoldsteinIO(C.Structure):
    _fields_ = [(n, D) for n in (
        "tstar_ocn", "sstar_ocn", "ustar_ocn", "vstar_ocn", "albedo_ocn", "go_ts", "go_u", "go_rho", "go_cost",
        "go_psi", "test_energy_ocean", "test_water_ocean")]


# every symbol include/cgenie_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cg_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(P)]),
    "cg_set_member_param": (C.c_int, [P, C.c_char_p, D]),
    "cg_initialise": (C.c_int, [P]),
    "cg_destroy": (C.c_int, [P]),
    "cg_last_error": (C.c_char_p, []),
    "cg_surflux_step": (C.c_int, [P, C.c_int, C.POINTER(SurfluxIO)]),
    "cg_embm_step": (C.c_int, [P, C.c_int, C.POINTER(EmbmIO)]),
    "cg_seaice_step": (C.c_int, [P, C.c_int, C.POINTER(SeaiceIO)]),
    "cg_goldstein_step": (C.c_int, [P, C.c_int, C.POINTER(GoldsteinIO)]),
    "cg_goldstein_mldta": (C.c_int, [P, C.c_int, D]),
    "cg_biogem_forcing": (C.c_int, [P, C.c_int64]),
    "cg_biogem_step": (C.c_int, [P, C.c_double, C.c_int64]),
    "cg_biogem_tracercoupling": (C.c_int, [P, D, D]),
    "cg_biogem_climate": (C.c_int, [P]),
    "cg_biogem_init_ocn": (C.c_int, [P]),
    "cg_biogem_climate_sol": (C.c_int, [P]),
    "cg_cpl_flux_ocnatm": (C.c_int, [P]),
    "cg_cpl_flux_ocnsed": (C.c_int, [P, C.c_double]),
    "cg_cpl_comp_ocnsed": (C.c_int, [P, C.c_int, C.c_int, C.c_int]),
    "cg_reinit_flux_rokocn": (C.c_int, [P]),
    "cg_exchange_begin_upload": (C.c_int, [P, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]),
    "cg_exchange_commit_upload": (C.c_int, [P, C.c_char_p, C.c_char_p]),
    "cg_exchange_begin_download": (C.c_int, [P, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]),
    "cg_exchange_wait": (C.c_int, [P]),
    "cg_biogem_sig_update": (C.c_int, [P, C.c_double, C.c_double]),
    "cg_biogem_sig_extended": (C.c_int, [P]),
    "cg_biogem_slice_update": (C.c_int, [P, C.c_double]),
    "cg_biogem_slice_reset": (C.c_int, [P]),
    "cg_biogem_sig_reset": (C.c_int, [P]),
    "cg_series_last_error": (C.c_char_p, []),
    "cg_biogem_series_write": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, STRS, I32, I32, C.c_int, STRS, I32, I32,
                                          D, C.c_int]),
    "cg_biogem_series_write_ext": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, STRS, I32, I32, C.c_int,
                                             STRS, I32, I32, D, D, C.c_double, C.c_double, C.c_int]),
    "cg_set_koverall": (C.c_int, [P, C.c_int64]),
    "cg_refresh_rho": (C.c_int, [P, C.c_int]),
    "cg_atchem_step": (C.c_int, [P, C.c_double]),
    "cg_run": (C.c_int, [P, C.c_int64]),
    "cg_field_size": (C.c_int64, [P, C.c_char_p]),
    "cg_sync_to_host": (C.c_int, [P, C.c_char_p, C.c_int, D, C.c_int64]),
    "cg_sync_from_host": (C.c_int, [P, C.c_char_p, C.c_int, D, C.c_int64]),
    "cg_sync_all_to_host": (C.c_int, [P, C.c_char_p, D, C.c_int64]),
    "cg_sync_all_from_host": (C.c_int, [P, C.c_char_p, D, C.c_int64]),
    "cg_wet_size": (C.c_int64, [P, C.c_char_p]),
    "cg_sync_all_wet_to_host": (C.c_int, [P, C.c_char_p, D, C.c_int64]),
    "cg_sync_all_wet_from_host": (C.c_int, [P, C.c_char_p, D, C.c_int64]),
    "cg_const_size": (C.c_int64, [P, C.c_char_p]),
    "cg_get_const": (C.c_int, [P, C.c_char_p, C.c_int, D, C.c_int64]),
    "cg_get_iconst": (C.c_int, [P, C.c_char_p, C.POINTER(C.c_int32), C.c_int64]),
    "cg_get_dims": (C.c_int, [P, C.POINTER(C.c_int32)]),
    "cg_global_means": (C.c_int, [P, D]),
    "cg_health": (C.c_int, [P, C.POINTER(C.c_int32)]),
    "cg_synchronize": (C.c_int, [P]),
    "cg_launch_count": (C.c_int64, [P, C.c_int]),
    "cg_timer_start": (C.c_int, [P]),
    "cg_timer_stop_ms": (C.c_int, [P, D]),
    "cg_profile_enable": (C.c_int, [P, C.c_int]),
    "cg_profile_get": (C.c_int, [P, C.c_char_p, D, C.POINTER(C.c_int64)]),
    "cg_set_tracer_variant": (C.c_int, [P, C.c_int]),
    "cg_tracer_variant_active": (C.c_int, [P]),
    "cg_set_biogem_fusion": (C.c_int, [P, C.c_int]),
    "cg_set_graphs": (C.c_int, [P, C.c_int]),
    "cg_tracer_create": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_int32), C.c_double, C.c_double, C.c_int, C.POINTER(P)]),
    "cg_tracer_set": (C.c_int, [P, D, D, D]),
    "cg_tracer_step": (C.c_int, [P, C.c_int]),
    "cg_tracer_get": (C.c_int, [P, D, D, D]),
    "cg_restart_last_error": (C.c_char_p, []),
    "cg_restart_goldstein_write": (C.c_int, [C.c_char_p] + [C.c_int] * 4 + [I32] + [D] * 8 + [I32]),
    "cg_restart_goldstein_read": (C.c_int, [C.c_char_p] + [C.c_int] * 4 + [D] * 5 + [I32]),
    "cg_restart_embm_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, D, D, D, I32]),
    "cg_restart_embm_read": (C.c_int, [C.c_char_p, C.c_int, C.c_int, D, I32]),
    "cg_restart_seaice_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int, I32, D, D, D, D, D, I32]),
    "cg_restart_seaice_read": (C.c_int, [C.c_char_p, C.c_int, C.c_int, D, D, D, I32]),
    "cg_restart_date_write": (C.c_int, [C.c_char_p, I32]),
    "cg_restart_date_read": (C.c_int, [C.c_char_p, I32]),
    "cg_restart_biogem_write": (C.c_int, [C.c_char_p] + [C.c_int] * 3 + [I32] + [D] * 6 + [C.c_int, STRS, STRS, D] * 2 +
                                [C.c_double, C.c_char_p]),
    "cg_slice_biogem_write_3d": (C.c_int, [C.c_char_p] + [C.c_int] * 3 + [I32] + [D] * 6 + [C.c_int, STRS, STRS, STRS, D, I32, I32, D] +
                                 [C.c_int, STRS, I32, I32, D] + [C.c_int, STRS, D] * 2 + [D, C.c_double, C.c_double, C.c_char_p]),
    "cg_restart_biogem_read": (C.c_int, [C.c_char_p] + [C.c_int] * 3 + [I32] + [C.c_int, STRS, D, I32] * 2),
    "cg_restart_atchem_write": (C.c_int, [C.c_char_p, C.c_int, C.c_int] + [D] * 4 + [C.c_int, STRS, STRS, D, C.c_double, C.c_char_p]),
    "cg_restart_atchem_read": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, STRS, D, I32]),
    "cg_restart_atchem_write_bin": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, I32, D]),
    "cg_restart_atchem_read_bin": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, I32, D, I32]),
    "cg_restart_biogem_write_bin": (C.c_int, [C.c_char_p] + [C.c_int] * 3 + [C.c_int, I32, D] * 2),
    "cg_restart_biogem_read_bin": (C.c_int, [C.c_char_p] + [C.c_int] * 3 + [C.c_int, I32, D, I32] * 2),
}


def load():
    """Load the compiled library or raise: the product has no CPU fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libcgenie_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cgenie_b200`); the B200 path has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB
