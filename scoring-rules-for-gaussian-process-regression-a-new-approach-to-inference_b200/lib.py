"""ctypes binding of libgpscore.so (include/gpscore.h).  No torch types cross this file.

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  If it is missing,
or no CUDA device is present, every entry point fails loudly: there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpscore.so")

GPS_OK, GPS_EINVAL, GPS_ECUDA, GPS_ENOTPD, GPS_ENODEVICE, GPS_ENOMEM, GPS_ESTATE = range(7)
GPS_CRPS, GPS_LOGS, GPS_NLML, GPS_DSS, GPS_KC = 0, 1, 2, 3, 4
SCORES = {"crps": GPS_CRPS, "logs": GPS_LOGS, "nlml": GPS_NLML, "dss": GPS_DSS, "kc": GPS_KC}
GRID_KINDS = {"nlml": 0, "crps": 1, "wrong_crps": 2, "logs": 3}

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> (restype, argtypes): exactly the declarations of include/gpscore.h
SIGNATURES = {
    "gps_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gps_destroy": (None, [_vp]),
    "gps_last_error": (C.c_char_p, [_vp]),
    "gps_version": (C.c_char_p, []),
    "gps_launch_count": (_i64, [_vp]),
    "gps_set_stream": (C.c_int, [_vp, _vp]),
    "gps_set_gemm_timing": (C.c_int, [_vp, C.c_int]),
    "gps_last_gemm_ms": (C.c_int, [_vp, _dp, C.POINTER(_i64)]),
    "gps_last_stage_ms": (C.c_int, [_vp, _dp]),
    "gps_set_data": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int]),
    "gps_full_eval": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp]),
    "gps_full_loo": (C.c_int, [_vp, _vp, _vp]),
    "gps_full_predict": (C.c_int, [_vp, _dp, _vp, _i64, _vp, _vp]),
    "gps_fitc_eval": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp]),
    "gps_fitc_acc_len": (C.c_int, [C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "gps_fitc_begin": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_double, C.c_int, _i64]),
    "gps_fitc_pass1": (C.c_int, [_vp, _vp]),
    "gps_fitc_pass2": (C.c_int, [_vp, _vp, _vp]),
    "gps_fitc_pass3": (C.c_int, [_vp, _vp, _vp]),
    "gps_fitc_finish": (C.c_int, [_vp, _vp, _vp, _dp, _dp, _dp]),
    "gps_comm_set_library": (C.c_int, [C.c_char_p]),
    "gps_comm_unique_id": (C.c_int, [_vp]),
    "gps_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "gps_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gps_comm_destroy": (C.c_int, [_vp]),
    "gps_comm_set_transport": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "gps_comm_allreduce_sum": (C.c_int, [_vp, _vp, _i64]),
    "gps_fitc_eval_sharded": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_double, C.c_int, _i64, _dp, _dp, _dp]),
    "gps_fitc_descend_sharded": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_double, C.c_int, _i64, C.c_double, C.c_double, C.c_int, _dp]),
    "gps_fitc_loo": (C.c_int, [_vp, _vp, _vp]),
    "gps_fitc_predict": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "gps_full_descend": (C.c_int, [_vp, _dp, C.c_int, C.c_double, C.c_int, _dp]),
    "gps_fitc_descend": (C.c_int, [_vp, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, _dp]),
    "gps_test_metrics": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double, _dp]),
    "gps_ard": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, C.c_double, _dp, C.c_int, _vp]),
    "gps_chol_solve": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "gps_matmul": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "gps_score": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_int, _dp]),
    "gps_grid_eval": (C.c_int, [_vp, _vp, _vp, C.c_int, _dp, _dp, _i64, C.c_int, _dp]),
}

# include/gpscore_debug.h (stage-level test hooks)
DEBUG_SIGNATURES = {
    "gps_dbg_gemm": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _i64, _i64, _i64, C.c_double, C.c_double, _vp, C.c_int]),
    "gps_dbg_factor": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "gps_dbg_fp64_peak": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "gps_dbg_set_variant": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gps_dbg_potf2_phases": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "gps_dbg_trace": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "gps_dbg_gram": (C.c_int, [_vp, _dp, _vp]),
    "gps_dbg_fused_phases": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "gps_dbg_launch_floor": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
}

_lib = None


class GpsError(RuntimeError):
    """Any non-zero return of the C-ABI.  `NotPositiveDefinite` keeps the reference's convention
    that a failed Cholesky raises RuntimeError (KF:726, K20:784)."""

    def __init__(self, code, msg):
        super().__init__("gpscore error %d: %s" % (code, msg))
        self.code = code


class NotPositiveDefinite(GpsError):
    pass


def load():
    """dlopen the in-tree library and attach the prototypes; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libgpscore.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "— the CUDA extension is required, there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, code):
    if code == GPS_OK:
        return
    msg = load().gps_last_error(ctx)
    msg = msg.decode() if msg else ""
    if code == GPS_ENOTPD:
        raise NotPositiveDefinite(code, msg)
    if code == GPS_ENODEVICE:
        raise GpsError(code, "no CUDA device: gpscore has no CPU fallback")
    raise GpsError(code, msg)
