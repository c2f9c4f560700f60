"""ctypes binding of libslamb200.so -- the same C ABI a C++ host links (include/slamb200.h).

No compute happens in Python and nothing here falls back to the CPU: if the shared object is
missing or no sm_100 device is present, the calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libslamb200.so")

DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                   ("distance", "<f4")])

# every symbol include/slamb200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _sz, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_int64
_pp = ctypes.POINTER(ctypes.c_void_p)
SYMBOLS = {
    "slamb200_init": (_i, [_i, _pp]),
    "slamb200_shutdown": (_i, [_vp]),
    "slamb200_version": (_i, []),
    "slamb200_last_error": (ctypes.c_char_p, []),
    "slamb200_launch_count": (_i64, [_vp]),
    "slamb200_synchronize": (_i, [_vp]),
    "slamb200_upload_desc": (_i, [_vp, _i, _vp, _i, _sz, _pp]),
    "slamb200_upload_desc_device": (_i, [_vp, _i, _vp, _i, _sz, _vp, _pp]),
    "slamb200_upload_desc_pinned": (_i, [_vp, _i, _vp, _i, _sz, _pp]),
    "slamb200_upload_desc_packed": (_i, [_vp, _i, _vp, _i, _sz, _pp]),
    "slamb200_set_pack_threads": (_i, [_vp, _i]),
    "slamb200_upload_desc_shared": (_i, [_vp, _i, _vp, _i, _sz, _pp]),
    "slamb200_desc_export": (_i, [_vp, _vp, _vp]),
    "slamb200_desc_import": (_i, [_vp, _vp, _pp]),
    "slamb200_desc_localize": (_i, [_vp, _vp, _pp]),
    "slamb200_free_desc": (_i, [_vp, _vp]),
    "slamb200_desc_rows": (_i, [_vp]),
    "slamb200_desc_kind": (_i, [_vp]),
    "slamb200_desc_exact_mode": (_i, [_vp]),
    "slamb200_knn2": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "slamb200_match_pair": (_i, [_vp, _i, _vp, _vp, _d, _vp, _i, _vp]),
    "slamb200_match_batch": (_i, [_vp, _i, _vp, _vp, _i, _d, _vp, _i, _vp]),
    "slamb200_match_batch_host": (_i, [_vp, _i, _vp, _i, _sz, _vp, _vp, _vp, _i, _d, _vp, _i, _vp]),
    "slamb200_match_window": (_i, [_vp, _i, _vp, _i, _d, _vp, _i, _vp]),
    "slamb200_match_batch_enqueue": (_i, [_vp, _i, _vp, _vp, _i, _d, _vp]),
    "slamb200_batch_fetch": (_i, [_vp, _vp, _i, _vp, _vp]),
    "slamb200_score_essential": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _d, _vp, _vp, _vp, _vp]),
    "slamb200_score_essential_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _d, _vp, _vp, _vp]),
    "slamb200_orb_compute": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _i, _vp, _vp, _vp, _vp]),
    "slamb200_sift_compute": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _i, _vp, _pp]),
    "slamb200_fast_detect": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _vp, _i, _vp]),
    "slamb200_fast_orb_compute": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _pp]),
    "slamb200_triangulate": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "slamb200_score_pnp": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _d, _i, _vp, _vp, _vp, _vp]),
    "slamb200_score_pnp_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _d, _i, _vp, _vp, _vp]),
    "slamb200_upload_pts": (_i, [_vp, _vp, _i, _sz, _pp]),
    "slamb200_free_pts": (_i, [_vp, _vp]),
    "slamb200_score_batch_enqueue": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _d, _vp]),
    "slamb200_batch_scores_fetch": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "slamb200_set_init": (_i, [_i, _pp]),
    "slamb200_set_shutdown": (_i, [_vp]),
    "slamb200_set_devices": (_i, [_vp]),
    "slamb200_set_ctx": (_vp, [_vp, _i]),
    "slamb200_set_owner": (_i, [_vp, _i, _i]),
    "slamb200_set_upload": (_i, [_vp, _i, _i, _vp, _i, _sz, _pp]),
    "slamb200_set_free_desc": (_i, [_vp, _vp]),
    "slamb200_mdesc_rows": (_i, [_vp]),
    "slamb200_set_match_batch": (_i, [_vp, _i, _vp, _vp, _i, _d, _vp, _i, _vp]),
    "slamb200_set_match_batch_enqueue": (_i, [_vp, _i, _vp, _vp, _i, _d]),
    "slamb200_set_batch_fetch": (_i, [_vp, _vp, _i, _vp, _vp]),
    "slamb200_profile_enable": (_i, [_vp, _i]),
    "slamb200_profile_enable_kinds": (_i, [_vp, ctypes.c_uint]),
    "slamb200_profile_read": (_i, [_vp, _vp, _vp]),
}

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_MATCHER, ERR_KIND, ERR_INTERNAL = 0, -1, -2, -3, -4, -5, -6
SIFT_BF, SIFT_FLANN, ORB_BF = 0, 1, 2
DESC_F32X128, DESC_U8X32 = 0, 1

_lib = None


class Slamb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libslamb200 error {code}: {msg}")
        self.code = code


def load():
    """Loads the shared object and types every exported entry point.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m slam_indoor_code_b200.build` "
                "(there is no Python/CPU fallback for the hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise Slamb200Error(rc, load().slamb200_last_error().decode(errors="replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
