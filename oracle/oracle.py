"""numpy front-end of the CPU oracle (oracle/mvsgi_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package
mvs_gi_b200 never does.  Parity status: pinned against vectors generated from
the reference's own modules (oracle/gen_golden.py -> tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from mvs_gi_b200 import abi  # POD struct definitions only

_DIR = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _DIR / "libmvsgi_oracle.so"
    src = _DIR / "mvsgi_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B", "libmvsgi_oracle.so"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _bstride(a, per_frame_elems):
    """Batch stride in elements; 0 when the array has a leading dim of 1 (shared)."""
    return 0 if a.shape[0] == 1 else per_frame_elems


def num_threads() -> int:
    return int(lib().orc_num_threads())


def bilinear_grid_sample(im, grid):
    """backports.py:11-86.  im [N,C,H,W], grid [N,Hg,Wg,2] -> [N,C,Hg,Wg]."""
    im, grid = _f32(im), _f32(grid)
    N, Cc, H, W = im.shape
    _, Hg, Wg, _ = grid.shape
    out = np.empty((N, Cc, Hg, Wg), np.float32)
    lib().orc_bilinear_grid_sample(_p(im), N, Cc, H, W, _p(grid), Hg, Wg, _p(out))
    return out


def sweep_std(feats, grids, grid_masks, masks, return_valid=False):
    """spherical_sweep_avg.py:38-136.
    feats [B,CAM,C,Hf,Wf]; grids [B|1,CAM,D,Ho,Wo,2]; grid_masks [B|1,CAM,D,Ho,Wo,1|-]; masks [B|1,CAM,1,Hm,Wm]."""
    feats, grids, masks = _f32(feats), _f32(grids), _f32(masks)
    gm = np.ascontiguousarray(np.asarray(grid_masks).reshape(grids.shape[:5]), dtype=np.uint8)
    B, CAM, Cc, Hf, Wf = feats.shape
    _, _, D, Ho, Wo, _ = grids.shape
    Hm, Wm = masks.shape[-2:]
    vol = np.empty((B, Cc, D, Ho, Wo), np.float32)
    valid = np.empty((B, D, Ho, Wo), np.uint8)
    lib().orc_sweep_std(
        _p(feats), _p(masks), _p(grids), _p(gm), B, CAM, Cc, D, Hf, Wf, Hm, Wm, Ho, Wo,
        C.c_int64(_bstride(masks, CAM * Hm * Wm)), C.c_int64(_bstride(grids, CAM * D * Ho * Wo * 2)),
        C.c_int64(_bstride(gm, CAM * D * Ho * Wo)), _p(vol), _p(valid))
    return (vol, valid.astype(bool)) if return_valid else vol


def sweep_concat(feats, grids):
    """spherical_sweep.py:38-68."""
    feats, grids = _f32(feats), _f32(grids)
    B, CAM, Cc, Hf, Wf = feats.shape
    _, _, D, Ho, Wo, _ = grids.shape
    vol = np.empty((B, CAM * Cc, D, Ho, Wo), np.float32)
    lib().orc_sweep_concat(_p(feats), _p(grids), B, CAM, Cc, D, Hf, Wf, Ho, Wo,
                           C.c_int64(_bstride(grids, CAM * D * Ho * Wo * 2)), _p(vol))
    return vol


def interp_out_size(n_in: int, scale_factor: float) -> int:
    """F.interpolate output size for a given scale_factor: floor(n * s)."""
    return int(np.floor(np.float64(n_in) * scale_factor))


def softargmin(costs, inv_dist_idx, interp_scale_factor=0.0, pre_interp=False):
    """distance_regressor.py:51-80.  costs [B,1,D,H,W] -> (inv_dist [B,1,Ho,Wo], norm_costs [B,D,Ho,Wo], argmax)."""
    costs = _f32(costs)
    B, _, D, H, W = costs.shape
    c0 = costs[:, 0]
    idx = _f32(np.asarray(inv_dist_idx).reshape(-1))
    interp = bool(pre_interp and interp_scale_factor > 0)
    Ho, Wo = (interp_out_size(H, interp_scale_factor), interp_out_size(W, interp_scale_factor)) if interp else (H, W)
    scale = np.float32(1.0 / interp_scale_factor) if interp else np.float32(1.0)
    inv = np.empty((B, 1, Ho, Wo), np.float32)
    nc = np.empty((B, D, Ho, Wo), np.float32)
    am = np.empty((B, Ho, Wo), np.int32)
    es = c0.strides
    lib().orc_softargmin(_p(c0), C.c_int64(es[0] // 4), C.c_int64(es[1] // 4), C.c_int64(es[2] // 4),
                         C.c_int64(es[3] // 4), B, D, H, W, Ho, Wo, int(interp), C.c_float(scale), C.c_float(scale),
                         _p(idx), _p(inv), _p(nc), _p(am))
    return inv, nc, am


def sweep_std_backward(feats, grids, grid_masks, masks, grad_vol):
    """d sum(grad_vol * sweep_std(feats, ...)) / d feats  (autograd of spherical_sweep_avg.py:38-136)."""
    feats, grids, masks, grad_vol = _f32(feats), _f32(grids), _f32(masks), _f32(grad_vol)
    gm = np.ascontiguousarray(np.asarray(grid_masks).reshape(grids.shape[:5]), dtype=np.uint8)
    B, CAM, Cc, Hf, Wf = feats.shape
    _, _, D, Ho, Wo, _ = grids.shape
    Hm, Wm = masks.shape[-2:]
    assert grad_vol.shape == (B, Cc, D, Ho, Wo)
    gf = np.empty_like(feats)
    lib().orc_sweep_std_backward(
        _p(feats), _p(masks), _p(grids), _p(gm), B, CAM, Cc, D, Hf, Wf, Hm, Wm, Ho, Wo,
        C.c_int64(_bstride(masks, CAM * Hm * Wm)), C.c_int64(_bstride(grids, CAM * D * Ho * Wo * 2)),
        C.c_int64(_bstride(gm, CAM * D * Ho * Wo)), _p(grad_vol), _p(gf))
    return gf


def sweep_concat_backward(feats_shape, grids, grad_vol):
    """adjoint of sweep_concat (spherical_sweep.py:38-68): grad_vol [B,CAM*C,D,Ho,Wo] -> grad_feats [B,CAM,C,Hf,Wf]."""
    grids, grad_vol = _f32(grids), _f32(grad_vol)
    B, CAM, Cc, Hf, Wf = feats_shape
    _, _, D, Ho, Wo, _ = grids.shape
    assert grad_vol.shape == (B, CAM * Cc, D, Ho, Wo)
    gf = np.empty(feats_shape, np.float32)
    lib().orc_sweep_concat_backward(_p(grids), B, CAM, Cc, D, Hf, Wf, Ho, Wo,
                                    C.c_int64(_bstride(grids, CAM * D * Ho * Wo * 2)), _p(grad_vol), _p(gf))
    return gf


def softargmin_backward(costs, inv_dist_idx, interp_scale_factor, pre_interp, g_inv=None, g_nc=None):
    """Backward of softargmin wrt costs[:, 0] -> grad [B,1,D,H,W]."""
    costs = _f32(costs)
    B, _, D, H, W = costs.shape
    c0 = costs[:, 0]
    idx = _f32(np.asarray(inv_dist_idx).reshape(-1))
    interp = bool(pre_interp and interp_scale_factor > 0)
    Ho, Wo = (interp_out_size(H, interp_scale_factor), interp_out_size(W, interp_scale_factor)) if interp else (H, W)
    scale = np.float32(1.0 / interp_scale_factor) if interp else np.float32(1.0)
    g_inv = None if g_inv is None else _f32(g_inv)
    g_nc = None if g_nc is None else _f32(g_nc)
    out = np.empty((B, D, H, W), np.float32)
    es = c0.strides
    lib().orc_softargmin_backward(_p(c0), C.c_int64(es[0] // 4), C.c_int64(es[1] // 4), C.c_int64(es[2] // 4),
                                  C.c_int64(es[3] // 4), B, D, H, W, Ho, Wo, int(interp), C.c_float(scale),
                                  C.c_float(scale), _p(idx), _p(g_inv), _p(g_nc), _p(out))
    return out[:, None]


def make_grids(rig_struct: abi.MvsgiRig):
    """create_grids_from_rays (multi_view_camera_model_dataset.py:530-557) in fp32 reference op order."""
    CAM, D, Ho, Wo = rig_struct.num_cams, rig_struct.num_cands, rig_struct.cv_height, rig_struct.cv_width
    grids = np.empty((CAM, D, Ho, Wo, 2), np.float32)
    gm = np.empty((CAM, D, Ho, Wo), np.uint8)
    lib().orc_make_grids(C.byref(rig_struct), _p(grids), _p(gm))
    return grids, gm.astype(bool)


def make_grids_f64(rig_struct: abi.MvsgiRig):
    CAM, D, Ho, Wo = rig_struct.num_cams, rig_struct.num_cands, rig_struct.cv_height, rig_struct.cv_width
    grids = np.empty((CAM, D, Ho, Wo, 2), np.float64)
    gm = np.empty((CAM, D, Ho, Wo), np.uint8)
    lib().orc_make_grids_f64(C.byref(rig_struct), _p(grids), _p(gm))
    return grids, gm.astype(bool)


def bf16_round(a: np.ndarray) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even) -> float32: what MVSGI_FLAG_FEATS_BF16 stores in the staged image
    (test infrastructure for the stated-tolerance mode, SURVEY.md §8c criterion 4)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def resample(rig_struct: abi.MvsgiRig, img, invalid_value: float = 0.0, grid=None, grid_valid=None,
             interpolation: str = "linear"):
    """CameraModelRotation (camera_model_sampler.py:55-120): img [N,C,H,W] -> (out [N,C,Ht,Wt], valid [Ht,Wt] bool).
    rig_struct as built by mvs_gi_b200.resample.resample_rig.  `grid` [Ht,Wt,2] + `grid_valid` [Ht,Wt]: sample with
    this grid (the CUDA kernel's own, from mvsgi_project_only) instead of the one evaluated here."""
    img = _f32(img)
    N, Cc, H, W = img.shape
    Ht, Wt = rig_struct.cv_height, rig_struct.cv_width
    out = np.empty((N, Cc, Ht, Wt), np.float32)
    valid = np.empty((Ht, Wt), np.uint8)
    if grid is not None:
        grid = _f32(grid).reshape(Ht, Wt, 2)
        grid_valid = np.ascontiguousarray(np.asarray(grid_valid).reshape(Ht, Wt), dtype=np.uint8)
    fn = {"linear": lib().orc_resample, "nearest": lib().orc_resample_nearest}[interpolation]
    fn(C.byref(rig_struct), _p(img), N, Cc, H, W, C.c_float(invalid_value), _p(out), _p(valid), _p(grid), _p(grid_valid))
    return out, valid.astype(bool)


def resample_blended(rig_struct: abi.MvsgiRig, img, blend_factor, invalid_value: float = 0.0, grid=None, grid_valid=None):
    """CameraModelRotation.blend_interpolation (camera_model_sampler.py:122-163): f * nearest + (1 - f) * linear with the
    blend factor image f = blend_func(img) itself sampled 'nearest'; invalid rays get `invalid_value`."""
    lin, valid = resample(rig_struct, img, 0.0, grid, grid_valid, "linear")
    near, _ = resample(rig_struct, img, 0.0, grid, grid_valid, "nearest")
    f, _ = resample(rig_struct, blend_factor, 0.0, grid, grid_valid, "nearest")
    out = (f * near + (np.float32(1) - f) * lin).astype(np.float32)
    out[:, :, ~valid] = np.float32(invalid_value)
    return out, valid
