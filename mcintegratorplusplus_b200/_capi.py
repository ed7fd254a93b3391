"""ctypes binding of include/mcig.h (libmcig.so). No torch, no numpy dependency beyond array marshalling.

The library is the product: if it is missing this module raises at import time with the build command — there is no
Python/CPU fallback for any compute entry point.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcig.so")

ERR_NAMES = {1: "invalid_argument", 2: "domain_error", 3: "runtime_error", 4: "cuda_error"}

ALLREDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.c_void_p)

# every symbol include/mcig.h declares: (restype, argtypes)
_dp, _ip, _u64p, _i64p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)
_ctx = C.c_void_p
SIGNATURES = {
    "mcig_last_error": (C.c_char_p, []),
    "mcig_version": (C.c_int, []),
    "mcig_device_count": (C.c_int, []),
    "mcig_register_plugin": (C.c_int, [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mcig_lookup_plugin": (C.c_int, [C.c_int, C.c_char_p]),
    "mcig_create": (_ctx, [C.c_int]),
    "mcig_destroy": (None, [_ctx]),
    "mcig_set_device": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_seed": (C.c_int, [_ctx, C.c_uint64]),
    "mcig_set_walker_seeds": (C.c_int, [_ctx, _u64p, C.c_int64]),
    "mcig_set_stream_position": (C.c_int, [_ctx, C.c_uint64]),
    "mcig_get_stream_position": (C.c_uint64, [_ctx]),
    "mcig_set_rng_mode": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_walkers": (C.c_int, [_ctx, C.c_int64, C.c_int64, C.c_int64]),
    "mcig_get_walkers": (C.c_int64, [_ctx]),
    "mcig_set_x": (C.c_int, [_ctx, _dp]),
    "mcig_set_x_walkers": (C.c_int, [_ctx, _dp]),
    "mcig_get_x": (C.c_int, [_ctx, C.c_int64, _dp]),
    "mcig_set_domain_unbound": (C.c_int, [_ctx]),
    "mcig_set_domain_ortho": (C.c_int, [_ctx, _dp, _dp]),
    "mcig_set_domain_plugin": (C.c_int, [_ctx, C.c_int, _dp, C.c_int, _dp, C.c_double]),
    "mcig_set_move": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, _ip]),
    "mcig_set_srrd_params": (C.c_int, [_ctx, C.c_int, _dp]),
    "mcig_set_move_plugin": (C.c_int, [_ctx, C.c_int, _dp, C.c_int, C.c_int, _ip]),
    "mcig_multistep_config": (C.c_int, [_ctx, C.c_int]),
    "mcig_multistep_add_pdf": (C.c_int, [_ctx, C.c_int, _dp, C.c_int]),
    "mcig_get_nsteps_sizes": (C.c_int, [_ctx]),
    "mcig_set_step": (C.c_int, [_ctx, C.c_int, C.c_double]),
    "mcig_get_step": (C.c_double, [_ctx, C.c_int]),
    "mcig_add_pdf": (C.c_int, [_ctx, C.c_int, _dp, C.c_int]),
    "mcig_pop_pdf": (C.c_int, [_ctx]),
    "mcig_clear_pdfs": (C.c_int, [_ctx]),
    "mcig_add_obs": (C.c_int, [_ctx, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mcig_set_callback": (C.c_int, [_ctx, C.c_int, _dp, C.c_int, C.c_int64]),
    "mcig_clear_callback": (C.c_int, [_ctx]),
    "mcig_get_callback_buffer": (C.c_int, [_ctx, _dp, C.c_int64]),
    "mcig_pop_obs": (C.c_int, [_ctx]),
    "mcig_clear_obs": (C.c_int, [_ctx]),
    "mcig_get_nobsdim": (C.c_int, [_ctx]),
    "mcig_set_autotune": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_double]),
    "mcig_set_allreduce": (C.c_int, [_ctx, ALLREDUCE_FN, C.c_void_p]),
    "mcig_integrate": (C.c_int, [_ctx, C.c_int64, _dp, _dp, C.c_int, C.c_int]),
    "mcig_get_acceptance_rate": (C.c_double, [_ctx]),
    "mcig_get_walker_results": (C.c_int, [_ctx, _dp, _dp]),
    "mcig_get_sums": (C.c_int, [_ctx, _dp, C.c_int]),
    "mcig_get_result_nobsdim": (C.c_int, [_ctx]),
    "mcig_get_cross_walker_error": (C.c_int, [_ctx, _dp, C.c_int]),
    "mcig_set_keep_samples": (C.c_int, [_ctx, C.c_int]),
    "mcig_comm_init_env": (C.c_int, []),
    "mcig_comm_get_unique_id": (C.c_int, [C.c_void_p]),
    "mcig_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mcig_comm_rank": (C.c_int, []),
    "mcig_comm_size": (C.c_int, []),
    "mcig_comm_finalize": (C.c_int, []),
    "mcig_attach_comm": (C.c_int, [_ctx, C.c_int]),
    "mcig_get_nstore": (C.c_int64, [_ctx, C.c_int]),
    "mcig_get_obs_data": (C.c_int, [_ctx, C.c_int, C.c_int64, _dp]),
    "mcig_get_timings": (C.c_int, [_ctx, _dp, _dp, _dp, _i64p]),
    "mcig_get_phase_timings": (C.c_int, [_ctx, _dp, _dp, _dp]),
    "mcig_estimate": (C.c_int, [C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp]),
    "mcig_estimate_blocks": (C.c_int, [C.c_int64, C.c_int, _dp, C.c_int64, _dp, _dp]),
    "mcig_set_block_size": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_state_placement": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_dynamic_scheduling": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_philox_rounds": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_lazy_accumulation": (C.c_int, [_ctx, C.c_int]),
    "mcig_set_device_calibration": (C.c_int, [_ctx, C.c_int]),
    "mcig_get_calibration_iterations": (C.c_int, [_ctx]),
    "mcig_get_decorrelation_chunks": (C.c_int, [_ctx]),
    "mcig_get_staging_chunks": (C.c_int64, [_ctx]),
    "mcig_store_on_file": (C.c_int, [_ctx, C.c_int, C.c_char_p, C.c_int]),
    "mcig_prebuild": (C.c_int, [_ctx]),
    "mcig_warmup": (C.c_int, [_ctx, C.c_int64, C.c_int, C.c_int]),
    "mcig_get_kernel_source": (C.c_int64, [_ctx, C.c_char_p, C.c_int64]),
    "mcig_measure_peaks": (C.c_int, [C.c_int, _dp, _dp]),
    "mcig_measure_philox_peak": (C.c_int, [C.c_int, _dp]),
}


class McigError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, "error %d" % code), msg))
        self.code = code
        self.msg = msg


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libmcig.so is not built: run `python -m mcintegratorplusplus_b200.build` (needs nvcc). "
                              "There is no CPU fallback for the sampling path.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(_lib, name)
            f.restype = res
            f.argtypes = args
    return _lib


def check(rc):
    if rc != 0:
        raise McigError(rc, lib().mcig_last_error().decode(errors="replace"))
