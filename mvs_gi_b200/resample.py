"""Fisheye -> surrogate-ERP image resampling on the GPU: the step before the feature extractor.

Drop-in for the way the reference's `CameraModelRotation` sampler is used at inference
(dsta_mvs/image_sampler/camera_model_sampler.py:12-120, driven per camera and per frame by
api/inference_pytorch.py:61-74): same constructor meaning (raw model, target model, R_raw_fisheye, default invalid
value), same call signature `sampler(img, interpolation='linear', invalid_pixel_value=None) -> (sampled, valid_mask)`,
same result.  The sampling grid is not precomputed and stored: the kernel evaluates the projection from ~200 bytes of
calibration (`csrc/resample.cuh`), and a whole batch of images that share the camera model and orientation is
resampled in one launch (the reference asserts batch size 1 and loops over cameras in Python).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import abi, ops
from . import calibration as cal


def resample_rig(raw_camera, target_camera: cal.EquirectCamera, R_raw_fisheye) -> cal.Rig:
    """The calibration the resample kernel takes, expressed as a one-camera, one-candidate rig: the target model plays
    the cost-volume camera, the raw model is camera 0 with the *direct* rotation R_raw_fisheye
    (x_raw = R_raw_fisheye @ x_target, camera_model_sampler.py:26-29) and no translation."""
    R = np.asarray(R_raw_fisheye, dtype=np.float32).reshape(3, 3)
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R.T.astype(np.float64)  # Rig stores T_cv_cam and hands the kernel its inverse rotation = R
    return cal.Rig(target_camera, [raw_camera], [T], [1.0])


class CameraModelRotation:
    """B200 counterpart of the reference sampler of the same name (camera_model_sampler.py:12)."""

    def __init__(self, camera_model_raw, camera_model_target: cal.EquirectCamera, R_raw_fisheye,
                 default_invalid_value: float = 0.0):
        if not isinstance(camera_model_target, cal.EquirectCamera):
            raise TypeError("the target of the resampler is the Equirectangular surrogate "
                            "(configs/dataset/surrogate_models_*.yaml)")
        if isinstance(camera_model_raw, cal.EquirectCamera):
            # an Equirectangular RAW image is sampled with border padding in the reference
            # (padding_mode_if_being_sampled = 'border', camera_models.py:633); the kernel pads with zeros
            raise TypeError("the resampler takes fisheye raw models (DoubleSphere, LinearSphere); an Equirectangular raw "
                            "image needs border padding, which is not implemented")
        self.camera_model_raw = camera_model_raw
        self.camera_model_target = camera_model_target
        self.default_invalid_value = default_invalid_value
        self._rig = resample_rig(camera_model_raw, camera_model_target, R_raw_fisheye)
        self._struct = self._rig.to_struct()
        self._valid: Optional[Tensor] = None

    @property
    def shape(self) -> Tuple[int, int]:
        return self.camera_model_target.H, self.camera_model_target.W

    def __call__(self, img: Tensor, interpolation: str = "linear", invalid_pixel_value=None, blend_func=None):
        """img [N, C, H, W] or [C, H, W] with any strides — e.g. `frames.permute(0, 3, 1, 2)` of the HWC camera buffers,
        no copy — float32, or uint8 which is scaled to [0, 1] on load like preprocess_imgs does
        (api/inference_pytorch.py:55-57), on a CUDA device -> (sampled [N, C, Ht, Wt] float32, valid_mask [Ht, Wt] bool).
        interpolation: 'linear' (what preprocess_imgs uses), 'nearest', or 'blended' with `blend_func(img) -> factor in
        [0, 1]` (1 = nearest only), the three modes of the reference sampler (camera_model_sampler.py:98-163)."""
        if interpolation == "blended":
            return self.blend_interpolation(img, blend_func, invalid_pixel_value)
        if interpolation not in ("linear", "nearest"):
            raise NotImplementedError(f"mvs_gi_b200 resampler: interpolation {interpolation!r} is not one of "
                                      "'linear', 'nearest', 'blended'")
        if blend_func is not None:
            raise ValueError("blend_func is only used by interpolation='blended'")
        invalid = float(self.default_invalid_value if invalid_pixel_value is None else invalid_pixel_value)
        return self._sample(self._as_batch(img), interpolation == "nearest", invalid)

    def blend_interpolation(self, img: Tensor, blend_func, invalid_pixel_value=None):
        """f * nearest + (1 - f) * linear, f = blend_func(img) sampled 'nearest' (camera_model_sampler.py:122-163)."""
        if blend_func is None:
            raise ValueError("interpolation='blended' needs blend_func")
        invalid = float(self.default_invalid_value if invalid_pixel_value is None else invalid_pixel_value)
        img = self._as_batch(img)
        if img.dtype == torch.uint8:  # the reference blends the converted float image (convert_input)
            img = img.float() / 255.0
        lin, valid = self._sample(img, False, 0.0)
        near, _ = self._sample(img, True, 0.0)
        f, _ = self._sample(self._as_batch(blend_func(img).to(torch.float32)), True, 0.0)
        out = f * near + (1 - f) * lin
        out[..., ~valid] = invalid
        return out, valid

    def _as_batch(self, img: Tensor) -> Tensor:
        if not img.is_cuda:
            raise RuntimeError(f"mvs_gi_b200: `img` must be a CUDA tensor (got {img.device}); there is no CPU path")
        if img.dim() == 3:
            img = img.unsqueeze(0)
        if img.dim() != 4:
            raise RuntimeError("img must be [N, C, H, W] or [C, H, W]")
        if img.dtype not in (torch.float32, torch.uint8):
            raise RuntimeError(f"img dtype {img.dtype} not supported (float32, uint8)")
        return img

    def _sample(self, img: Tensor, nearest: bool, invalid: float):
        N, Cc, H, W = img.shape
        raw = self.camera_model_raw
        if (H, W) != (raw.H, raw.W):  # check_input_shape, camera_model_sampler.py:92-96
            raise RuntimeError(f"Wrong input image shape. Expect {(raw.H, raw.W)}, got {(H, W)}")
        Ht, Wt = self.shape
        lib = abi.load_library()
        out = torch.empty((N, Cc, Ht, Wt), dtype=torch.float32, device=img.device)
        valid = torch.empty((Ht, Wt), dtype=torch.uint8, device=img.device)
        a = abi.MvsgiResampleArgs()
        a.struct_size = C.sizeof(abi.MvsgiResampleArgs)
        a.flags = abi.RESAMPLE_IMG_U8 if img.dtype == torch.uint8 else 0  # v / 255 on load, no float copy of the image
        if nearest:
            a.flags |= abi.RESAMPLE_NEAREST
        a.batch, a.channels, a.in_h, a.in_w = N, Cc, H, W
        a.img = img.data_ptr()
        for i, s in enumerate(img.stride()):
            a.img_stride[i] = int(s)
        a.rig = C.pointer(self._struct)
        a.invalid_value = invalid
        a.out = out.data_ptr()
        a.valid_out = valid.data_ptr()
        with torch.cuda.device(img.device):
            abi.check(lib.mvsgi_resample_f32(C.byref(a), torch.cuda.current_stream(img.device).cuda_stream),
                      "mvsgi_resample_f32")
        return out, valid.bool()
