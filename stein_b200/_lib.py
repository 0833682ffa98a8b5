"""ctypes binding of libstein_b200.so (the C ABI declared in include/stein_b200.h).

There is no CPU implementation behind this module: if the shared library is
missing, or no B200 is visible, every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstein_b200.so")

STEIN_OK = 0
PHI_AUTO, PHI_DENSE_SIMT, PHI_FLASH_TC, PHI_FLASH_TC2, PHI_FLASH_TC3, PHI_FLASH_TC4, PHI_FLASH_TC5 = 0, 1, 2, 3, 4, 5, 6
OPT_ADAM, OPT_ADAGRAD = 0, 1
MEDIAN_AUTO, MEDIAN_FFMA, MEDIAN_TC, MEDIAN_TC1 = 0, 1, 2, 3

c_i64, c_i32, c_u32, c_u64 = ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64
c_f32, c_f64, c_vp = ctypes.c_float, ctypes.c_double, ctypes.c_void_p

HOOK = ctypes.CFUNCTYPE(ctypes.c_int, c_vp, c_vp, c_i64)
HOOK_GATHER = ctypes.CFUNCTYPE(ctypes.c_int, c_vp, c_vp, c_vp, c_i64)
HOOK_GATHER_ON = ctypes.CFUNCTYPE(ctypes.c_int, c_vp, c_vp, c_vp, c_i64, c_vp)


class SteinComm(ctypes.Structure):
    _fields_ = [("rank", c_i32), ("world", c_i32), ("user", c_vp),
                ("allreduce_sum_u64", HOOK), ("allreduce_sum_f64", HOOK),
                ("allgather_f32", HOOK_GATHER), ("allgather_f32_on", HOOK_GATHER_ON)]


class SteinLibraryError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/stein_b200.h one to one
SIGNATURES = {
    "stein_abi_version": (ctypes.c_int, []),
    "stein_ctx_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, c_vp]),
    "stein_ctx_destroy": (ctypes.c_int, [c_vp]),
    "stein_ctx_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "stein_ctx_set_comm": (ctypes.c_int, [c_vp, ctypes.POINTER(SteinComm)]),
    "stein_nccl_unique_id": (ctypes.c_int, [c_vp]),
    "stein_ctx_init_nccl": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "stein_ctx_set_phi_impl": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_ctx_set_median_impl": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_last_error": (ctypes.c_char_p, [c_vp]),
    "stein_ctx_launch_count": (c_i64, [c_vp]),
    "stein_ctx_phi_route": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i32), ctypes.POINTER(c_f32), ctypes.POINTER(c_f32)]),
    "stein_ctx_set_phi_guard_tol": (ctypes.c_int, [c_vp, c_f32]),
    "stein_ctx_profile_enable": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_ctx_profile_read": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_f64), ctypes.POINTER(c_i64)]),
    "stein_ld": (c_i64, [c_i64]),
    "stein_rows_padded": (c_i64, [c_i64]),
    "stein_row_norms": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "stein_sqdist_hist": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64,
                                         c_u32, c_u32, c_u32, c_vp]),
    "stein_num_tiles": (c_i64, [c_i64]),
    "stein_tile_coords": (ctypes.c_int, [c_i64, c_i64, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "stein_median_sqdist": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64,
                                           ctypes.POINTER(c_f32), ctypes.POINTER(c_f32),
                                           ctypes.POINTER(c_i32)]),
    "stein_median_values": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_f32)]),
    "stein_bandwidth": (c_f32, [c_f32, c_i64]),
    "stein_median_narrow": (ctypes.c_int, [ctypes.POINTER(c_u64), c_u32, c_u32, c_u32, c_u64,
                                           ctypes.POINTER(c_u32), ctypes.POINTER(c_u32),
                                           ctypes.POINTER(c_u32), ctypes.POINTER(c_u32)]),
    "stein_float_to_key": (c_u32, [c_f32]),
    "stein_key_to_float": (c_f32, [c_u32]),
    "stein_phi_workspace_bytes": (c_i64, [c_vp, c_i64, c_i64, c_i64]),
    "stein_phi": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32,
                                 c_vp, c_i64, c_vp, c_vp]),
    "stein_kernel_and_grad": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_vp,
                                             c_i64, c_vp, c_vp, c_i64]),
    "stein_imq_kernel_and_grad": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_vp,
                                                 c_i64, c_vp, c_vp, c_i64]),
    "stein_phi_from_kernel": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64,
                                             c_vp, c_vp]),
    "stein_clip_adam_step": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_f64, c_f64,
                                            c_f64, c_i64]),
    "stein_clip_adagrad_step": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_f64, c_f64,
                                               c_i64]),
    "stein_score_linear": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "stein_score_logistic": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_f64,
                                            c_f64, c_f64, c_vp]),
    "stein_score_bnn": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64,
                                       c_f64, c_f64, c_f64, c_vp]),
    "stein_score_gaussian_mixture": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_f64, c_vp]),
    "stein_predict_linear": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "stein_predict_bnn": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "stein_engine_create": (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, c_i64, c_i64, ctypes.c_int,
                                           c_f64, c_f64, c_f64, c_f64]),
    "stein_engine_destroy": (ctypes.c_int, [c_vp]),
    "stein_engine_local_rows": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "stein_engine_buffers": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                            ctypes.POINTER(c_vp), ctypes.POINTER(c_i64),
                                            ctypes.POINTER(c_i64)]),
    "stein_engine_set_particles": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "stein_engine_get_particles": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "stein_engine_set_scores": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "stein_engine_get_phi": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int]),
    "stein_engine_step": (ctypes.c_int, [c_vp]),
    "stein_engine_phi_only": (ctypes.c_int, [c_vp]),
    "stein_engine_apply_phi": (ctypes.c_int, [c_vp]),
    "stein_engine_sumsq_dev": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "stein_engine_set_hyper": (ctypes.c_int, [c_vp, c_f64, c_f64, c_f64, c_f64]),
    "stein_engine_update_particles_host": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int]),
    "stein_ctx_trace_enable": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_ctx_trace_read": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_i64]),
    "stein_engine_set_prefetch": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_engine_set_device_bandwidth": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "stein_engine_device_bandwidth_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "stein_engine_prefetch_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "stein_engine_particles_changed": (ctypes.c_int, [c_vp]),
    "stein_engine_ipc_handle": (ctypes.c_int, [c_vp, c_vp]),
    "stein_engine_set_peer_handles": (ctypes.c_int, [c_vp, c_vp]),
    "stein_engine_set_bandwidth": (ctypes.c_int, [c_vp, c_f32]),
    "stein_engine_last": (ctypes.c_int, [c_vp, ctypes.POINTER(c_f32), ctypes.POINTER(c_f32),
                                         ctypes.POINTER(c_f64), ctypes.POINTER(c_i32)]),
    "stein_engine_get_state": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_f64),
                                              c_vp, c_vp, ctypes.c_int]),
    "stein_engine_set_state": (ctypes.c_int, [c_vp, c_i64, c_f64, c_vp, c_vp, ctypes.c_int]),
}

_lib = None


def load():
    """Load libstein_b200.so; raises SteinLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SteinLibraryError(
            "%s is missing: build it with `python __graft_entry__.py build` "
            "(make -C stein_b200/csrc). stein_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.stein_abi_version() != 2:
        raise SteinLibraryError("libstein_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != STEIN_OK:
        msg = load().stein_last_error(ctx)
        raise SteinLibraryError("libstein_b200 error %d: %s" % (rc, (msg or b"").decode()))
