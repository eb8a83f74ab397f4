"""ctypes mirror of include/mvsgi_b200.h and the loader for libmvsgi_b200.so.

The product path has exactly one implementation: the CUDA library.  If the
library is missing this module raises at first use — there is no CPU or eager
PyTorch fallback (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

ABI_VERSION = 13
MAX_CAMS = 4
MAX_CANDS = 64
MAX_PEERS = 7
P2P_HANDLE_BYTES = 64

CAM_EQUIRECT, CAM_DOUBLE_SPHERE, CAM_LINEAR_SPHERE = 0, 1, 2
GRID_SUPPLIED, GRID_PROJECT = 0, 1
AGG_STD, AGG_CONCAT = 0, 1
MASK_F32, MASK_U8, MASK_PACKED = 0, 1, 2
TAP_TABLE_OFF, TAP_TABLE_BUILD, TAP_TABLE_USE = 0, 1, 2
FLAG_FEATS_BF16 = 1
FLAG_STAGE_ONLY = 2
FLAG_SWEEP_ONLY = 4
FLAG_VOL_CHANNELS_LAST = 8
FLAG_CTA_TABLE = 16
FLAG_STAGED_FEATS = 32
FLAG_VOL_BF16 = 64
# MVSGI_E_* (negative return codes; positive ones are cudaError_t)
E_NULL, E_SHAPE, E_STRIDE, E_UNSUPPORTED, E_WORKSPACE, E_VERSION = -1, -2, -3, -4, -5, -6
RESAMPLE_IMG_U8 = 1
RESAMPLE_NEAREST = 2

LIB_NAME = "libmvsgi_b200.so"
_PKG_DIR = Path(__file__).resolve().parent


class MvsgiCamera(C.Structure):
    _fields_ = [
        ("model", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("reserved", C.c_int32),
        ("rot", C.c_float * 9),
        ("trans", C.c_float * 3),
        ("p", C.c_float * 12),
    ]


class MvsgiRig(C.Structure):
    _fields_ = [
        ("num_cams", C.c_int32),
        ("num_cands", C.c_int32),
        ("cv_width", C.c_int32),
        ("cv_height", C.c_int32),
        ("cv_lon0", C.c_float),
        ("cv_lon_span_pixel", C.c_float),
        ("cv_lat0", C.c_float),
        ("cv_lat_range", C.c_float),
        ("cams", MvsgiCamera * MAX_CAMS),
        ("dist", C.c_float * MAX_CANDS),
    ]


class MvsgiSweepArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("aggregate", C.c_int32),
        ("grid_mode", C.c_int32),
        ("flags", C.c_uint32),
        ("batch", C.c_int32),
        ("num_cams", C.c_int32),
        ("feat_chs", C.c_int32),
        ("num_cands", C.c_int32),
        ("feat_h", C.c_int32),
        ("feat_w", C.c_int32),
        ("mask_h", C.c_int32),
        ("mask_w", C.c_int32),
        ("out_h", C.c_int32),
        ("out_w", C.c_int32),
        ("feats", C.c_void_p),
        ("feats_stride", C.c_int64 * 5),
        ("masks", C.c_void_p),
        ("masks_dtype", C.c_int32),
        ("reserved0", C.c_int32),
        ("masks_stride", C.c_int64 * 4),
        ("grids", C.c_void_p),
        ("grids_stride", C.c_int64 * 5),
        ("grid_masks", C.c_void_p),
        ("grid_masks_stride", C.c_int64 * 5),
        ("rig", C.POINTER(MvsgiRig)),
        ("vol", C.c_void_p),
        ("valid_out", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("cand_offset", C.c_int32),
        ("vol_num_cands", C.c_int32),
        ("num_peers", C.c_int32),
        ("reserved1", C.c_int32),
        ("peer_vols", C.c_void_p * MAX_PEERS),
        ("octet_order", C.c_void_p),
        ("order_runs", C.c_int32),
        ("order_units", C.c_int32),
        ("cta_times_out", C.c_void_p),
        ("tap_table", C.c_void_p),
        ("tap_table_bytes", C.c_int64),
        ("tap_table_mode", C.c_int32),
        ("reserved2", C.c_int32),
    ]


class MvsgiProjectArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("reserved", C.c_int32),
        ("rig", C.POINTER(MvsgiRig)),
        ("grids", C.c_void_p),
        ("grid_masks", C.c_void_p),
    ]


class MvsgiResampleArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("batch", C.c_int32),
        ("channels", C.c_int32),
        ("in_h", C.c_int32),
        ("in_w", C.c_int32),
        ("img", C.c_void_p),
        ("img_stride", C.c_int64 * 4),
        ("rig", C.POINTER(MvsgiRig)),
        ("invalid_value", C.c_float),
        ("reserved", C.c_int32),
        ("out", C.c_void_p),
        ("valid_out", C.c_void_p),
    ]


class MvsgiSoftArgminArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("batch", C.c_int32),
        ("num_cands", C.c_int32),
        ("in_h", C.c_int32),
        ("in_w", C.c_int32),
        ("out_h", C.c_int32),
        ("out_w", C.c_int32),
        ("src_scale_h", C.c_float),
        ("src_scale_w", C.c_float),
        ("costs", C.c_void_p),
        ("costs_stride", C.c_int64 * 4),
        ("inv_dist_idx", C.c_void_p),
        ("inv_dist", C.c_void_p),
        ("norm_costs", C.c_void_p),
        ("argmax", C.c_void_p),
    ]


class MvsgiSweepBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("fwd", C.POINTER(MvsgiSweepArgs)),
        ("grad_vol", C.c_void_p),
        ("grad_feats", C.c_void_p),
        ("grad_feats_stride", C.c_int64 * 5),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
    ]


class MvsgiSoftArgminBwdArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("flags", C.c_uint32),
        ("fwd", C.POINTER(MvsgiSoftArgminArgs)),
        ("grad_inv_dist", C.c_void_p),
        ("grad_norm_costs", C.c_void_p),
        ("grad_costs", C.c_void_p),
        ("grad_costs_stride", C.c_int64 * 4),
    ]


# Every symbol include/mvsgi_b200.h declares: name -> (restype, argtypes).
EXPORTS = {
    "mvsgi_sweep_workspace_bytes": (C.c_int64, [C.POINTER(MvsgiSweepArgs)]),
    "mvsgi_sweep_f32": (C.c_int, [C.POINTER(MvsgiSweepArgs), C.c_void_p]),
    "mvsgi_sweep_launch_count": (C.c_int, [C.POINTER(MvsgiSweepArgs)]),
    "mvsgi_sweep_balance_shape": (C.c_int, [C.POINTER(MvsgiSweepArgs), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mvsgi_sweep_octet_cost": (C.c_int, [C.POINTER(MvsgiSweepArgs), C.c_void_p, C.c_void_p]),
    "mvsgi_sweep_grid": (C.c_int, [C.POINTER(MvsgiSweepArgs)]),
    "mvsgi_sweep_tap_table_bytes": (C.c_int64, [C.POINTER(MvsgiSweepArgs)]),
    "mvsgi_project_only_f32": (C.c_int, [C.POINTER(MvsgiProjectArgs), C.c_void_p]),
    "mvsgi_resample_f32": (C.c_int, [C.POINTER(MvsgiResampleArgs), C.c_void_p]),
    "mvsgi_mask_pitch_words": (C.c_int32, [C.c_int32]),
    "mvsgi_pack_masks": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mvsgi_softargmin_f32": (C.c_int, [C.POINTER(MvsgiSoftArgminArgs), C.c_void_p]),
    "mvsgi_sweep_backward_workspace_bytes": (C.c_int64, [C.POINTER(MvsgiSweepBwdArgs)]),
    "mvsgi_sweep_backward_f32": (C.c_int, [C.POINTER(MvsgiSweepBwdArgs), C.c_void_p]),
    "mvsgi_softargmin_backward_f32": (C.c_int, [C.POINTER(MvsgiSoftArgminBwdArgs), C.c_void_p]),
    "mvsgi_p2p_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "mvsgi_p2p_free": (C.c_int, [C.c_void_p]),
    "mvsgi_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mvsgi_p2p_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mvsgi_p2p_close": (C.c_int, [C.c_void_p]),
    "mvsgi_strerror": (C.c_char_p, [C.c_int]),
    "mvsgi_abi_version": (C.c_int, []),
    "mvsgi_launch_counter": (C.c_int64, []),
}

_lib = None


def library_path() -> Path:
    override = os.environ.get("MVSGI_B200_LIB")
    return Path(override) if override else _PKG_DIR / LIB_NAME


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C mvs_gi_b200/csrc`). mvs_gi_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.mvsgi_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {got}, python side expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load_library().mvsgi_strerror(code)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else code} (code {code})")
