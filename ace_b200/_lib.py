"""ctypes binding of libace_b200.so (the C ABI declared in include/ace_b200.h).

There is no CPU or PyTorch fallback: if the shared library is absent (and cannot be built
because nvcc is missing) every entry point raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libace_b200.so")

_lock = threading.Lock()
_lib = None


class AceError(RuntimeError):
    pass


class SfnoConfig(ctypes.Structure):
    _fields_ = [
        ("img_h", ctypes.c_int), ("img_w", ctypes.c_int),
        ("in_chans", ctypes.c_int), ("out_chans", ctypes.c_int),
        ("embed_dim", ctypes.c_int), ("num_layers", ctypes.c_int),
        ("lmax", ctypes.c_int), ("mmax", ctypes.c_int),
        ("mlp_hidden", ctypes.c_int),
        ("operator_type", ctypes.c_int),
        ("normalization", ctypes.c_int),
        ("pos_embed", ctypes.c_int),
        ("big_skip", ctypes.c_int),
        ("norm_eps", ctypes.c_float),
    ]


class StepConfig(ctypes.Structure):
    _fields_ = [
        ("n_in", ctypes.c_int), ("n_out", ctypes.c_int), ("n_prog", ctypes.c_int), ("n_forcing", ctypes.c_int),
        ("in_kind_host", ctypes.POINTER(ctypes.c_int)),
        ("in_index_host", ctypes.POINTER(ctypes.c_int)),
        ("out_prog_index_host", ctypes.POINTER(ctypes.c_int)),
        ("in_mean_host", ctypes.POINTER(ctypes.c_float)),
        ("in_std_host", ctypes.POINTER(ctypes.c_float)),
        ("out_mean_host", ctypes.POINTER(ctypes.c_float)),
        ("out_std_host", ctypes.POINTER(ctypes.c_float)),
        ("residual_prediction", ctypes.c_int),
        ("out_force_positive_host", ctypes.POINTER(ctypes.c_int)),
        ("ocean_out_index", ctypes.c_int),
        ("ocean_interpolate", ctypes.c_int),
    ]


class SlabOceanConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("prog_sst", "out_dlw_sfc", "out_ulw_sfc", "out_dsw_sfc", "out_usw_sfc", "out_lhf", "out_shf")] + [
        ("timestep_seconds", ctypes.c_double)]


class CsfnoConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "img_h", "img_w", "in_chans", "out_chans", "embed_dim", "num_layers", "lmax", "mmax", "mlp_hidden", "pos_embed", "big_skip",
        "normalize_big_skip", "affine_norms", "embed_dim_scalar", "embed_dim_labels", "embed_dim_noise", "embed_dim_pos")] + [
        ("norm_eps", ctypes.c_float), ("filter_residual", ctypes.c_int), ("filter_output", ctypes.c_int),
        ("clip_latent_global_means", ctypes.c_int)]


class CorrectorConfig(ctypes.Structure):
    _fields_ = [
        ("n_out", ctypes.c_int), ("n_prog", ctypes.c_int), ("nz", ctypes.c_int),
        ("hw", ctypes.c_longlong),
        ("area_weights_host", ctypes.POINTER(ctypes.c_float)),
        ("ak_host", ctypes.POINTER(ctypes.c_double)),
        ("bk_host", ctypes.POINTER(ctypes.c_double)),
        ("out_prog_index_host", ctypes.POINTER(ctypes.c_int)),
        ("out_ps", ctypes.c_int),
        ("out_wat_host", ctypes.POINTER(ctypes.c_int)),
        ("out_precip", ctypes.c_int), ("out_lhf", ctypes.c_int), ("out_adv", ctypes.c_int),
        ("prog_ps", ctypes.c_int),
        ("prog_wat_host", ctypes.POINTER(ctypes.c_int)),
        ("conserve_dry_air", ctypes.c_int),
        ("moisture_mode", ctypes.c_int),
        ("timestep_seconds", ctypes.c_double),
        ("zero_global_mean_moisture_advection", ctypes.c_int),
        ("out_frozen", ctypes.c_int),
        ("clip_frozen_precipitation", ctypes.c_int),
        ("energy_mode", ctypes.c_int),
        ("unaccounted_heating", ctypes.c_double),
        ("out_temp_host", ctypes.POINTER(ctypes.c_int)),
        ("prog_temp_host", ctypes.POINTER(ctypes.c_int)),
        ("n_forcing", ctypes.c_int), ("forcing_hgt", ctypes.c_int),
        ("out_dlw_sfc", ctypes.c_int), ("out_ulw_sfc", ctypes.c_int), ("out_dsw_sfc", ctypes.c_int), ("out_usw_sfc", ctypes.c_int),
        ("out_shf", ctypes.c_int), ("out_usw_toa", ctypes.c_int), ("out_ulw_toa", ctypes.c_int),
    ]


# every symbol include/ace_b200.h declares: (restype, argtypes)
_VP, _I, _LL, _CP = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_char_p
SIGNATURES = {
    "ace_version": (_I, []),
    "ace_last_error": (_CP, []),
    "ace_set_option": (_I, [_CP, _I]),
    "ace_get_option": (_I, [_CP]),
    "ace_launch_count": (_LL, []),
    "ace_profile_report": (_I, [ctypes.c_char_p, _I]),
    "ace_set_scope_callback": (_I, [_VP, _VP]),
    "ace_debug_scope": (_I, [_CP]),
    "ace_sht_plan_create": (_I, [_I, _I, _I, _I, _VP, _VP, ctypes.POINTER(_VP)]),
    "ace_sht_plan_destroy": (None, [_VP]),
    "ace_sht_forward": (_I, [_VP, _VP, _VP, _LL, _VP]),
    "ace_sht_inverse": (_I, [_VP, _VP, _VP, _LL, _VP]),
    "ace_sfno_create": (_I, [ctypes.POINTER(SfnoConfig), _VP, _VP, ctypes.POINTER(_VP)]),
    "ace_sfno_destroy": (None, [_VP]),
    "ace_sfno_set_param": (_I, [_VP, _CP, _VP, _LL, _VP]),
    "ace_sfno_finalize": (_I, [_VP]),
    "ace_sfno_forward": (_I, [_VP, _VP, _VP, _I, _VP]),
    "ace_sfno_query": (_I, [_VP, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_LL)]),
    "ace_stepper_create": (_I, [_VP, ctypes.POINTER(StepConfig), ctypes.POINTER(_VP)]),
    "ace_stepper_destroy": (None, [_VP]),
    "ace_stepper_step": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "ace_dev_gemm": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "ace_csfno_create": (_I, [ctypes.POINTER(CsfnoConfig), _VP, _VP, ctypes.POINTER(_VP)]),
    "ace_csfno_destroy": (None, [_VP]),
    "ace_csfno_set_param": (_I, [_VP, _CP, _VP, _LL, _VP]),
    "ace_csfno_query": (_I, [_VP, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_LL)]),
    "ace_stepper_create_conditional": (_I, [_VP, ctypes.POINTER(StepConfig), ctypes.POINTER(_VP)]),
    "ace_stepper_set_slab_ocean": (_I, [_VP, ctypes.POINTER(SlabOceanConfig)]),
    "ace_stepper_set_context": (_I, [_VP, _VP, _VP]),
    "ace_csfno_finalize": (_I, [_VP, _VP]),
    "ace_label_embed": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _VP]),
    "ace_label_pos_embed": (_I, [_VP, _VP, _VP, _I, _I, _LL, _VP, _VP]),
    "ace_csfno_forward": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "ace_isotropic_noise": (_I, [_VP, _VP, _VP, _VP, _VP, _LL, _VP]),
    "ace_corrector_create": (_I, [ctypes.POINTER(CorrectorConfig), ctypes.POINTER(_VP)]),
    "ace_corrector_destroy": (None, [_VP]),
    "ace_corrector_seed": (_I, [_VP, _VP, _I, _VP]),
    "ace_corrector_reset": (_I, [_VP]),
    "ace_corrector_is_seeded": (_I, [_VP]),
    "ace_corrector_get_state": (_I, [_VP, _VP, _I, _VP]),
    "ace_corrector_set_state": (_I, [_VP, _VP, _I, _VP]),
    "ace_corrector_needs_next": (_I, [_VP]),
    "ace_corrector_apply": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "ace_stepper_set_corrector": (_I, [_VP, _VP]),
    "ace_hpx_forward": (_I, [_VP, _I, _VP, _VP, _LL, _VP]),
    "ace_hpx_inverse": (_I, [_VP, _I, _VP, _VP, _LL, _VP]),
    "ace_weighted_moments": (_I, [_VP, _VP, _VP, _LL, _LL, _VP, _VP]),
    "ace_zonal_mean": (_I, [_VP, _LL, _I, _I, _VP, _VP]),
    "ace_time_sum": (_I, [_VP, _I, _LL, _I, _LL, _LL, _VP, _VP]),
    "ace_power_spectrum": (_I, [_VP, _LL, _I, _I, _VP, _VP]),
}


def load(build_if_missing=True):
    """Load (building first if the .so is missing and nvcc exists) and return the ctypes library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise AceError(f"{LIB_PATH} not found")
            from . import build as _build

            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise AceError(
                    f"libace_b200.so is missing and could not be built ({e}); ace_b200 has no CPU/PyTorch fallback"
                ) from e
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # raises AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load().ace_last_error()
        raise AceError(f"ace_b200 error {rc}: {msg.decode() if msg else '?'}")


def set_option(key, value):
    check(load().ace_set_option(key.encode(), int(value)))


def get_option(key):
    return load().ace_get_option(key.encode())


def launch_count():
    return int(load().ace_launch_count())


def profile_report():
    """{kernel name: (launch count, total milliseconds)} since the last call (option "profile" must be 1)."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = load().ace_profile_report(buf, len(buf))
    out = {}
    for line in buf.raw[:n].decode().splitlines():
        name, count, ms = line.rsplit(" ", 2)
        out[name] = (int(count), float(ms))
    return out


def current_stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
