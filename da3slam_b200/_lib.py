"""ctypes binding of libda3s.so (include/da3s.h).  There is NO CPU fallback: if the
library is missing or CUDA is unavailable every entry point raises."""
from __future__ import annotations

import ctypes as C
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DA3S_LIB") or os.path.join(HERE, "libda3s.so")       # DA3S_LIB: a variant build (build.py --out=)

OK, EINVAL, EALIGN, ENOMEM, ECUDA, ETOOFEW = 0, -1, -2, -3, -4, -5

# flags of da3s_unproject_filter
UNPROJ_CLOSED, UNPROJ_KINV, UNPROJ_FAST = 0x0, 0x1, 0x2
UNPROJ_WORLD, UNPROJ_OUT_F64 = 0x4, 0x8
MASK_CONF_GT, MASK_CONF_GE, MASK_CONF_FLOOR, MASK_DEPTH, MASK_WORLD_Z = 0x10, 0x20, 0x40, 0x80, 0x100
SIM3_PER_FRAME = 0x200
CAM_CLOSED_FORM, CAM_GENERAL_INV = 0, 1
SEL_VALUES, SEL_POSITIVE, SEL_RATIO = 0, 1, 2
SEL_MEDIAN, SEL_PERCENTILE = 0, 1
UMEYAMA_WEIGHTED, UMEYAMA_MEAN, UMEYAMA_NORMRATIO, UMEYAMA_LEGACY_TRACE = 0, 1, 2, 3
ICP_SIM3, ICP_RIGID = 0, 1
ROW_LEN = 16

CAM_BYTES = 8 * 4 + 21 * 8          # sizeof(da3s_cam) = 200
PAIR_BYTES = 6 * 8                  # sizeof(da3s_pair)
AUX_DOUBLES = 8


class AlignOpts(C.Structure):
    _fields_ = [
        ("world", C.c_int), ("depth_scale_mode", C.c_int), ("depth_conf_th", C.c_float), ("depth_eps", C.c_float),
        ("valid_depth", C.c_int), ("conf_thr_override", C.c_float), ("huber", C.c_int), ("huber_delta", C.c_double),
        ("max_iterations", C.c_int), ("tol", C.c_double), ("min_points", C.c_int), ("n_hyp", C.c_int),
        ("ransac_thr", C.c_float), ("ransac_min_inliers", C.c_int), ("precise", C.c_int),
    ]


class SelectSeg(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("ca", C.c_void_p), ("cb", C.c_void_p), ("n", C.c_longlong),
        ("kind", C.c_int), ("stat", C.c_int), ("percent", C.c_float), ("conf_th", C.c_float), ("eps", C.c_float),
        ("reserved", C.c_float),
    ]


class SelectOut(C.Structure):
    _fields_ = [("n_valid", C.c_longlong), ("lo", C.c_float), ("hi", C.c_float), ("value", C.c_float), ("gamma", C.c_float)]


class Pair(C.Structure):
    _fields_ = [("depth_a", C.c_void_p), ("conf_a", C.c_void_p), ("depth_b", C.c_void_p), ("conf_b", C.c_void_p),
                ("cam_a", C.c_void_p), ("cam_b", C.c_void_p)]


class FrameJob(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("conf", C.c_void_p), ("cam", C.c_void_p), ("sim3", C.c_void_p),
                ("conf_thr", C.c_void_p), ("xyz", C.c_void_p), ("mask", C.c_void_p)]


class VoxelJob(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("rgb", C.c_void_p), ("mask", C.c_void_p), ("n", C.c_longlong)]


class ExportJob(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("conf", C.c_void_p), ("cam", C.c_void_p), ("sim3", C.c_void_p),
                ("conf_thr", C.c_void_p), ("rgb", C.c_void_p)]


assert C.sizeof(FrameJob) == 56 and C.sizeof(VoxelJob) == 32 and C.sizeof(ExportJob) == 48
assert C.sizeof(Pair) == PAIR_BYTES and C.sizeof(SelectSeg) == 64 and C.sizeof(SelectOut) == 24

_P, _I, _L, _F, _D, _ULL = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_ulonglong

# every symbol include/da3s.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "da3s_create": (_I, [_I, C.c_size_t, C.POINTER(_P)]),
    "da3s_destroy": (_I, [_P]),
    "da3s_strerror": (C.c_char_p, [_I]),
    "da3s_version": (_I, []),
    "da3s_last_cuda_error": (_I, [_P]),
    "da3s_launch_count": (_ULL, [_P]),
    "da3s_enable_peer_access": (_I, [_P, _I]),
    "da3s_measure_fp32_peak": (_I, [_P, _I, C.POINTER(_D), _P]),
    "da3s_ransac_round_of": (_I, [_I, _I]),
    "da3s_kernel_timers": (_I, [_P, _I]),
    "da3s_kernel_time": (_I, [_P, _I, C.POINTER(_D), C.POINTER(_I), C.POINTER(_I), C.POINTER(_D)]),
    "da3s_build_cams": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "da3s_unproject_filter": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _F, _F, _P, _P, _P, _P, _P]),
    "da3s_unproject_filter_jobs": (_I, [_P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _P]),
    "da3s_apply_sim3": (_I, [_P, _P, _I, _L, _P, _P, _I, _P]),
    "da3s_select": (_I, [_P, _P, _I, _L, _P, _P]),
    "da3s_filter_points": (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _L, _P, _P, _P, _P]),
    "da3s_align_opts_default": (None, [C.POINTER(AlignOpts)]),
    "da3s_align_pairs": (_I, [_P, _P, _I, _I, _I, _I, C.POINTER(AlignOpts), _P, _P, _P, _P, _P]),
    "da3s_accumulate_sim3": (_I, [_P, _P, _I, _P, _P]),
    "da3s_pair_thresholds": (_I, [_P, _P, _I, _I, _I, _I, C.POINTER(AlignOpts), _P, _P, _P, _P]),
    "da3s_ransac_score": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _I, _F, _P, _P]),
    "da3s_ransac_hypotheses": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "da3s_ransac_inlier_mask": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _F, _P, _P]),
    "da3s_umeyama_points": (_I, [_P, _P, _P, _I, _P, _I, _L, _P, _P, _L, _I, _P, _P]),
    "da3s_irls_points": (_I, [_P, _P, _P, _I, _P, _P, _L, _P, _P, _L, _D, _I, _D, _P, _P]),
    "da3s_voxel_begin": (_I, [_P, _L, _P]),
    "da3s_voxel_insert": (_I, [_P, _P, _P, _P, _L, _F, _P]),
    "da3s_voxel_insert_jobs": (_I, [_P, _P, _I, _L, _I, _F, _P]),
    "da3s_icp_points": (_I, [_P, _P, _L, _P, _L, _I, _I, _D, _D, _I, _P, _P]),
    "da3s_voxel_send": (_I, [_P, _I, _I, _P, _P, _P, _ULL, _L, _P]),
    "da3s_voxel_merge_inbox": (_I, [_P, _P, _P, _P, _ULL, _I, _L, _P]),
    "da3s_unproject_voxel_jobs": (_I, [_P, _P, _I, _I, _I, _I, _F, _F, _F, _F, _P]),
    "da3s_voxel_finish": (_I, [_P, _F, _L, _P, _P, _P, _P, _P, _P, _P]),
    "da3s_align_pairs_host": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(AlignOpts), _P, _P, _P]),
}


class Da3sError(RuntimeError):
    def __init__(self, code, where, msg):
        super().__init__(f"{where}: {msg} (code {code})")
        self.code = code


_lib = None


def load():
    """dlopen libda3s.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m da3slam_b200.build` "
            "(the submap-alignment path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, where):
    if code != OK:
        lib = load()
        raise Da3sError(code, where, lib.da3s_strerror(code).decode())


def default_opts(**kw) -> AlignOpts:
    o = AlignOpts()
    load().da3s_align_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown align option {k!r}")
        setattr(o, k, v)
    return o


NAN = math.nan

# da3s_kernel_time ids (include/da3s.h)
TIMED_RANSAC_SCORE, TIMED_IRLS, TIMED_EXPORT_VOXEL, TIMED_VOXEL_EMIT = 0, 1, 2, 3
TIMED_NAMES = {TIMED_RANSAC_SCORE: "ransac_score_kernel", TIMED_IRLS: "pair_moments_mixed_kernel",
               TIMED_EXPORT_VOXEL: "export_voxel_kernel", TIMED_VOXEL_EMIT: "voxel_emit_kernel"}
