"""ctypes binding of libtdl.so (include/tdl.h).  There is no fallback: if the library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDL_LIB_PATH") or os.path.join(HERE, "libtdl.so")   # override: kernel experiments only

TDL_MAX_SRC = 4
TDL_MAX_SCALES = 4
TDL_ABI_VERSION = 4
TDL_LAYOUT_NCHW, TDL_LAYOUT_NHWC = 0, 1
TDL_DTYPE_F32, TDL_DTYPE_BF16 = 0, 1

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


class PhotoArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("S", C.c_int32), ("nscales", C.c_int32),
        ("disp_h", C.c_int32 * TDL_MAX_SCALES), ("disp_w", C.c_int32 * TDL_MAX_SCALES),
        ("automask", C.c_int32), ("disp_norm", C.c_int32), ("align_corners", C.c_int32), ("reserved0", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("photo_coef", C.c_float * TDL_MAX_SCALES), ("smooth_coef", C.c_float * TDL_MAX_SCALES),
        ("smooth_alpha", C.c_float), ("reserved1", C.c_float),
        ("noise_seed", C.c_uint64),
        ("target", _vp), ("src", _vp * TDL_MAX_SRC), ("disp", _vp * TDL_MAX_SCALES),
        ("P", _vp), ("invK", _vp),
        ("noise", (_vp * TDL_MAX_SRC) * TDL_MAX_SCALES),
        ("warped", (_vp * TDL_MAX_SRC) * TDL_MAX_SCALES),
        ("min_index", _vp * TDL_MAX_SCALES),
        ("workspace", _vp), ("workspace_bytes", C.c_uint64),
        ("losses", _vp),
        ("dlosses", _vp), ("d_disp", _vp * TDL_MAX_SCALES), ("dP", _vp),
    ]


class FeatArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("S", C.c_int32),
        ("disp_h", C.c_int32), ("disp_w", C.c_int32), ("align_corners", C.c_int32),
        ("layout", C.c_int32), ("dtype", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("coef", C.c_float), ("reserved0", C.c_float),
        ("tgt", _vp), ("src", _vp * TDL_MAX_SRC), ("disp", _vp), ("P", _vp), ("invK", _vp),
        ("warped", _vp * TDL_MAX_SRC), ("min_index", _vp),
        ("workspace", _vp), ("workspace_bytes", C.c_uint64),
        ("loss", _vp),
        ("dloss", _vp), ("d_tgt", _vp), ("d_src", _vp * TDL_MAX_SRC), ("d_disp", _vp), ("dP", _vp),
        ("bwd_scratch", _vp), ("bwd_scratch_bytes", C.c_uint64),
    ]


class EdgeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("alpha", C.c_float), ("first_coef", C.c_float), ("second_coef", C.c_float), ("reserved0", C.c_float),
        ("feature", _vp), ("image", _vp),
        ("workspace", _vp), ("workspace_bytes", C.c_uint64),
        ("loss", _vp), ("dloss", _vp), ("d_feature", _vp),
    ]


TDL_MAX_LEVELS = 5


class EdgeMultiArgs(C.Structure):
    _fields_ = [("nlevels", C.c_int32), ("reserved0", C.c_int32), ("level", EdgeArgs * TDL_MAX_LEVELS)]


class ReconArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("coef", C.c_float),
        ("pred", _vp), ("target", _vp), ("mask", _vp),
        ("workspace", _vp), ("workspace_bytes", C.c_uint64),
        ("loss", _vp), ("dloss", _vp), ("d_pred", _vp),
    ]


class PoseArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("invert", C.c_int32),
        ("axisangle", _vp), ("translation", _vp), ("T", _vp),
        ("dT", _vp), ("d_axisangle", _vp), ("d_translation", _vp),
    ]


class ProjArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32),
        ("K", _vp), ("inv_K", _vp), ("T", _vp * TDL_MAX_SRC),
        ("P_full", _vp), ("P_half", _vp), ("invK3", _vp), ("invKh3", _vp),
        ("dP_full", _vp), ("dP_half", _vp), ("dT", _vp * TDL_MAX_SRC),
    ]


class InputArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("nframes", C.c_int32),
        ("erase_count", C.c_int32), ("erase_h", C.c_int32), ("erase_w", C.c_int32), ("reserved0", C.c_int32),
        ("frames", _vp * (TDL_MAX_SRC + 1)),
        ("jitter", _vp), ("order", _vp), ("do_aug", _vp), ("do_flip", _vp), ("holes", _vp),
        ("color", _vp * (TDL_MAX_SRC + 1)), ("color_aug", _vp * (TDL_MAX_SRC + 1)), ("mask", _vp),
        ("workspace", _vp), ("workspace_bytes", C.c_uint64),
    ]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int32), ("total_ms", C.c_double)]


EXPORTS = ["tdl_abi_version", "tdl_strerror", "tdl_set_option", "tdl_get_option", "tdl_launch_count", "tdl_profile_begin", "tdl_profile_end",
           "tdl_photo_ws_bytes", "tdl_photo_fwd", "tdl_photo_bwd",
           "tdl_feat_ws_bytes", "tdl_feat_bwd_scratch_bytes", "tdl_feat_fwd", "tdl_feat_bwd",
           "tdl_edge_ws_bytes", "tdl_edge_smooth_fwd", "tdl_edge_smooth_bwd", "tdl_edge_smooth_multi_fwd", "tdl_edge_smooth_multi_bwd",
           "tdl_recon_ws_bytes", "tdl_recon_fwd", "tdl_recon_bwd", "tdl_pose_fwd", "tdl_pose_bwd",
           "tdl_proj_fwd", "tdl_proj_bwd", "tdl_input_ws_bytes", "tdl_input_fwd"]

_lib = None


class TdlError(RuntimeError):
    pass


def lib():
    """Loads libtdl.so once.  Raises (never falls back) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TdlError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                       "(nvcc, sm_100a); there is no CPU fallback for this path")
    L = C.CDLL(LIB_PATH)
    L.tdl_abi_version.restype = C.c_int
    L.tdl_strerror.restype = C.c_char_p
    L.tdl_strerror.argtypes = [C.c_int]
    L.tdl_launch_count.restype = C.c_int
    L.tdl_launch_count.argtypes = [C.c_char_p]
    L.tdl_set_option.restype = C.c_int
    L.tdl_set_option.argtypes = [C.c_char_p, C.c_int]
    L.tdl_get_option.restype = C.c_int
    L.tdl_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.tdl_profile_begin.restype = C.c_int
    L.tdl_profile_end.restype = C.c_int
    L.tdl_profile_end.argtypes = [C.POINTER(KernelTime), C.c_int]
    L.tdl_photo_ws_bytes.restype = C.c_uint64
    L.tdl_photo_ws_bytes.argtypes = [C.c_int32] * 5 + [C.POINTER(C.c_int32)] * 2
    L.tdl_feat_ws_bytes.restype = C.c_uint64
    L.tdl_feat_ws_bytes.argtypes = [C.c_int32] * 5
    L.tdl_feat_bwd_scratch_bytes.restype = C.c_uint64
    L.tdl_feat_bwd_scratch_bytes.argtypes = [C.c_int32] * 5
    L.tdl_recon_ws_bytes.restype = C.c_uint64
    L.tdl_recon_ws_bytes.argtypes = []
    L.tdl_input_ws_bytes.restype = C.c_uint64
    L.tdl_input_ws_bytes.argtypes = [C.c_int32] * 2
    L.tdl_edge_ws_bytes.restype = C.c_uint64
    L.tdl_edge_ws_bytes.argtypes = [C.c_int32] * 4
    for name, T in (("tdl_photo_fwd", PhotoArgs), ("tdl_photo_bwd", PhotoArgs),
                    ("tdl_feat_fwd", FeatArgs), ("tdl_feat_bwd", FeatArgs),
                    ("tdl_edge_smooth_fwd", EdgeArgs), ("tdl_edge_smooth_bwd", EdgeArgs),
                    ("tdl_edge_smooth_multi_fwd", EdgeMultiArgs), ("tdl_edge_smooth_multi_bwd", EdgeMultiArgs),
                    ("tdl_recon_fwd", ReconArgs), ("tdl_recon_bwd", ReconArgs),
                    ("tdl_pose_fwd", PoseArgs), ("tdl_pose_bwd", PoseArgs),
                    ("tdl_proj_fwd", ProjArgs), ("tdl_proj_bwd", ProjArgs), ("tdl_input_fwd", InputArgs)):
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(T), C.c_void_p]
    if L.tdl_abi_version() != TDL_ABI_VERSION:
        raise TdlError(f"libtdl.so ABI {L.tdl_abi_version()} != binding {TDL_ABI_VERSION}: rebuild")
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise TdlError(f"{what} failed ({rc}): {lib().tdl_strerror(rc).decode()}")


def set_option(name: str, value: int) -> None:
    """Process-wide kernel-selection switch (tests / experiments): include/tdl.h tdl_set_option."""
    check(lib().tdl_set_option(name.encode(), int(value)), f"tdl_set_option({name})")


def get_option(name: str) -> int:
    v = C.c_int(0)
    check(lib().tdl_get_option(name.encode(), C.byref(v)), f"tdl_get_option({name})")
    return v.value


class options:
    """with options(photo_sparse_max=0, no_tma=1): ...   -- restores the previous values on exit."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)


def launch_count(entry: str) -> int:
    return lib().tdl_launch_count(entry.encode())


def profile_begin():
    check(lib().tdl_profile_begin(), "tdl_profile_begin")


def profile_end():
    """-> {kernel name: (launches, total_ms)} measured with CUDA events on the launch stream."""
    rows = (KernelTime * 64)()
    n = lib().tdl_profile_end(rows, 64)
    return {rows[i].name.decode(): (rows[i].launches, rows[i].total_ms) for i in range(n)}
