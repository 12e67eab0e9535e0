"""ctypes binding of libkmpc.so (include/kmpc.h).  There is NO fallback: if the library is
missing or a call fails, an exception is raised."""
import ctypes
import os

from .build import LIB_PATH

_c = ctypes
_vp, _i, _i64, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double


class KmpcError(RuntimeError):
    pass


class LoopConfigC(_c.Structure):  # mirrors kmpc_loop_config
    _fields_ = [("S", _i64), ("nz", _i), ("n", _i), ("N", _i), ("out_mode", _i), ("out_row", _i),
                ("du_aug", _i), ("update", _i), ("rls_flags", _i), ("c_pairs_next", _i),
                ("skip_first_barx", _i), ("shared_model", _i), ("lift_kind", _i), ("lift_mode", _i),
                ("plant_kind", _i), ("rk4_variant", _i), ("first_post_step", _i), ("max_iter", _i),
                ("h", _d), ("q", _d), ("rw", _d), ("lb", _d), ("ub", _d), ("u_lb", _d), ("u_ub", _d),
                ("lam", _d), ("p0", _d), ("q0", _d), ("tol", _d), ("path", _i), ("qp_cold", _i)]


class LoopBuffersC(_c.Structure):  # mirrors kmpc_loop_buffers
    _fields_ = [("x", _vp), ("z", _vp), ("u_prev", _vp), ("A", _vp), ("B", _vp), ("C", _vp),
                ("KA", _vp), ("P", _vp), ("barX", _vp), ("barQ", _vp), ("r", _vp),
                ("params_pre", _vp), ("params_post", _vp), ("cx", _vp), ("log_x", _vp),
                ("log_u", _vp), ("status", _vp), ("log_capacity", _i64)]


_PROTOS = {
    "kmpc_strerror": (_c.c_char_p, [_i]),
    "kmpc_last_cuda_error": (_c.c_char_p, []),
    "kmpc_version": (_i, []),
    "kmpc_launch_count": (_i64, []),
    "kmpc_encoder_create": (_i, [_c.POINTER(_vp), _c.POINTER(_vp), _c.POINTER(_vp), _c.POINTER(_i), _i, _vp]),
    "kmpc_encoder_destroy": (_i, [_vp]),
    "kmpc_encoder_out_dim": (_i, [_vp, _i]),
    "kmpc_encode": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "kmpc_encode_ex": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
    "kmpc_encoder_has_tc": (_i, [_vp]),
    "kmpc_gram_from_snapshots_ex": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp]),
    "kmpc_gram_from_trajectories_ex": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "kmpc_rbf_lift": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "kmpc_gram_pack_len": (_i64, [_i, _i]),
    "kmpc_gram_accumulate": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    "kmpc_gram_from_snapshots": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _vp, _vp]),
    "kmpc_gram_from_trajectories": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _vp, _vp]),
    "kmpc_edmd_solve": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "kmpc_rls_update": (_i, [_vp] * 11 + [_i64, _i, _i, _d, _i, _vp]),
    "kmpc_qp_first_move": (_i, [_vp] * 8 + [_d, _d, _i, _i, _i, _i64, _i, _vp, _vp, _vp, _i, _d, _vp]),
    "kmpc_plant_step": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _d, _vp]),
    "kmpc_ctx_create": (_i, [_c.POINTER(_vp), _c.POINTER(LoopConfigC), _c.POINTER(LoopBuffersC), _vp, _i, _vp]),
    "kmpc_ctx_destroy": (_i, [_vp]),
    "kmpc_closed_loop_steps": (_i, [_vp, _i, _vp]),
    "kmpc_ctx_step_index": (_i64, [_vp]),
    "kmpc_ctx_reset": (_i, [_vp, _i, _vp]),
    "kmpc_ctx_is_fused": (_i, [_vp]),
    "kmpc_measure_fp64_peak": (_i, [_c.POINTER(_d), _c.POINTER(_d), _vp]),
    "kmpc_generate_snapshots": (_i, [_vp, _vp, _vp, _i, _i, _d, _i64, _i, _vp, _vp, _vp, _vp]),
    "kmpc_open_loop_predict": (_i, [_vp] * 6 + [_i, _i, _i64, _i, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    "kmpc_window_losses": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _i, _i64, _i64, _vp, _vp]),
    "kmpc_closed_loop_steps_timed": (_i, [_vp, _i, _vp, _c.POINTER(_c.c_float)]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)
_lib = None


def lib():
    """Load libkmpc.so (once).  Raises KmpcError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KmpcError("libkmpc.so not found at %s -- run `python -m koopman_online_updated_mpc_b200.build` "
                            "(there is no CPU fallback)" % LIB_PATH)
        handle = _c.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.kmpc_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.kmpc_last_cuda_error().decode()
        raise KmpcError("libkmpc call failed (%d): %s" % (rc, msg))


def launch_count():
    return int(lib().kmpc_launch_count())


def measure_fp64_peak():
    """(dmma_tflops, dfma_tflops) of the current GPU: the fp64 roofline denominators (synchronises)."""
    import torch
    a, b = _d(), _d()
    check(lib().kmpc_measure_fp64_peak(_c.byref(a), _c.byref(b), torch.cuda.current_stream().cuda_stream))
    return a.value, b.value
