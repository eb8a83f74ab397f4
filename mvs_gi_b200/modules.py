"""Drop-in nn.Modules for the reference's sweep / cv_builder / distance-regression components.

Same constructor arguments, attributes, method signatures, state_dict keys and
error behaviour as the reference classes they replace, so a Lightning model or
api/inference_pytorch.py can swap them in:

  SphericalSweepStdMasked   dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:9-167
  SphericalSweep            dsta_mvs/model/cost_volume_builder/spherical_sweep.py:9-102
  DistanceRegressorWithFixedCandidates
                            dsta_mvs/model/distance_regressor/distance_regressor.py:7-80

The sweep itself runs in the fused sm_100a kernel; `post_vol` (Conv3d + BatchNorm3d
+ LeakyReLU, dsta_mvs/model/common/common_modules.py:82-115) stays on cuDNN as in
the reference and keeps its parameter names (`post_vol.conv_layer.weight`,
`post_vol.norm_layer.*`).
"""
from __future__ import annotations

import math
import weakref
from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from . import abi, ops
from .calibration import Rig


class _NoOp(nn.Module):
    def forward(self, x):
        return x


# dsta_mvs/model/common/__init__.py:7-23
RELU_TYPE = {"original": nn.ReLU, "leaky": nn.LeakyReLU, "none": _NoOp}
NORM3D_TYPE = {"batch": nn.BatchNorm3d, "instance": nn.InstanceNorm3d, "none": _NoOp}


class _ConvBlk3d(nn.Module):
    """Conv3d -> norm -> activation with the reference's sub-module names (common_modules.py:82-115)."""

    def __init__(self, in_chs: int, out_chs: int, kernel_size: int, norm_layer: nn.Module, activation: nn.Module):
        super().__init__()
        self.conv_layer = nn.Conv3d(in_chs, out_chs, kernel_size, padding=kernel_size // 2, bias=False, stride=1)
        self.norm_layer = norm_layer
        self.activation = activation

    def forward(self, x: Tensor) -> Tensor:
        return self.activation(self.norm_layer(self.conv_layer(x)))


class _SweepBase(nn.Module):
    _aggregate = abi.AGG_STD

    def __init__(self, num_cams: int, feat_chs: int, post_k_sz: int, norm_type: str = "batch",
                 relu_type: str = "leaky"):
        super().__init__()
        self.num_cams = num_cams
        self.feat_chs = feat_chs
        self.norm_type = NORM3D_TYPE[norm_type]
        self.relu_type = RELU_TYPE[relu_type]
        self.grid_sample_mode = "bilinear"
        self.post_vol = _ConvBlk3d(self.feat_chs, self.feat_chs, post_k_sz, norm_layer=self.norm_type(self.feat_chs),
                                   activation=self.relu_type())
        # Fused in-kernel projection (north_star (b)): when a rig is attached, `grids`/`grid_masks` passed to
        # forward() are ignored and never read from HBM.
        self._rig: Optional[Rig] = None
        self._rig_blob: Optional[Tensor] = None
        self._static_masks: Optional[Tensor] = None
        # Packed mask bits (ops.pack_masks) of the masks tensor last seen, reused while the caller keeps passing the same
        # tensor object at the same version: inference (api/inference_pytorch.py holds one masks tensor on the device)
        # then reads no float masks and launches no packing kernel per call.
        self._mask_cache = None  # (weakref to the masks tensor, its _version, packed bits, (Hm, Wm))
        # True: every new masks tensor is checked for negative / NaN values when it is packed (one device->host sync per
        # new tensor; off by default so that a forward stays CUDA-graph capturable); set_static_masks always checks.
        self.check_masks = False
        # Cost-balanced work assignment (ops.balanced_octet_order): with an attached rig and static masks the per-octet
        # cost of the sweep is known in advance; the octets are dealt to the resident CTAs so that every CTA samples
        # about the same number of cameras.  Same volume bit for bit; computed once per shape (one synchronisation).
        # Measured on B200 (profiles/r3j_balance.jsonl): the per-CTA times of the round-robin deal spread from 332 to
        # 405 us at G16V and follow the count of (octet, candidate, camera loaded by the warp); dealt by that count they
        # lie within 345-353 us and the kernel ends 7 % earlier.  Set to False to keep the round-robin deal.
        self.balance_runs = True
        # ... and corrected by measurement (ops.calibrated_octet_order): CTA slots that run the kernel slower than the
        # rest get less work.  A dozen extra sweep launches at the first call with a given shape.
        self.calibrate_runs = True
        self.balance_report = None
        self._order_cache = None  # (key, int32 [runs, units] table or None, uint8 tap table or None)
        # Tap tables kept across calls (ops.build_tap_table, MvsgiSweepArgs.tap_table): with an attached rig and static
        # masks the kernel's (output offset, validity, tap offsets, bilinear fractions) records are the same at every
        # forward — what the grids the reference's dataset precomputes once per rig are to its sweep — so they are
        # built at the first call with a given shape and copied in (one cp.async.bulk per work item) afterwards.
        # Same volume bit for bit; pays when the batch does not amortise the build (single frames, high-D).
        # Costs records x (12 * cams + 12) bytes of device memory (G16V 18 MB, 64 candidates at 160x640: 288 MB).
        self.keep_tap_table = True
        # Memory layout of the volume the sweep kernel writes.  False = the reference's NCDHW.  True = the same
        # [B,C,D,Ho,Wo] tensor with torch.channels_last_3d strides: the kernel then stores 512 contiguous bytes per
        # warp and post_vol's Conv3d runs cuDNN's NDHWC path on it.  Values are identical.
        self.vol_channels_last = False
        # Stage the features as bfloat16 inside the sweep (fp32 arithmetic): stated-tolerance mode, 1e-2 * max|vol|.
        self.feats_bf16 = False
        # Write the volume as bfloat16 (channels-last) and run post_vol in bfloat16 on the tensor cores (row f2).
        self.vol_bf16 = False

    # ---- B200-specific configuration (not part of the reference interface) ----
    def attach_rig(self, rig: Optional[Rig]) -> "_SweepBase":
        """Evaluate the sampling grid in-kernel from this calibration instead of reading `grids`."""
        self._rig = rig
        self._rig_blob = ops.rig_to_blob(rig.to_struct()) if rig is not None else None
        self._order_cache = None
        return self

    def set_vol_channels_last(self, on: bool = True) -> "_SweepBase":
        """Emit the cost volume channels-last (NDHWC in memory) and keep post_vol's weights in the same format."""
        self.vol_channels_last = bool(on)
        self.post_vol.to(memory_format=torch.channels_last_3d if on else torch.contiguous_format)
        return self

    def set_vol_bf16(self, on: bool = True) -> "_SweepBase":
        """Stated-tolerance mode for the hand-off to post_vol: the sweep rounds its fp32 result to
        bfloat16 on the way out (|vol_bf16 - vol| <= 2^-8 |vol|, zero pattern unchanged), in the channels-last layout, and
        post_vol (Conv3d + BatchNorm3d + LeakyReLU, common_modules.py:82-115) runs in bfloat16 — half the volume bytes
        written and re-read, and cuDNN's tensor-core convolution instead of the fp32 one.  forward() returns float32
        like the reference (one cast of post_vol's output); sweep() returns the bfloat16 volume.  Implies
        set_vol_channels_last(True)."""
        self.vol_bf16 = bool(on)
        if on:
            self.set_vol_channels_last(True)
        self.post_vol.to(torch.bfloat16 if on else torch.float32)
        return self

    def set_feats_bf16(self, on: bool = True) -> "_SweepBase":
        """Stated-tolerance mode (SURVEY.md §8c criterion 4): taps read bfloat16-rounded features, half the L1/L2
        bytes; masks / validity are unchanged, |vol - fp32 vol| <= 1e-2 * max|vol|.  Forward only (no validity output);
        the backward differentiates the fp32 formulation (straight-through)."""
        self.feats_bf16 = bool(on)
        return self

    def set_static_masks(self, masks: Optional[Tensor]) -> "_SweepBase":
        """Inference: masks are the same for every frame ([1|B, CAM, 1, Hm, Wm]); only frame 0 is kept and shared."""
        self._static_masks = None if masks is None else masks[:1].contiguous()
        self._mask_cache = None
        self._order_cache = None
        if self._static_masks is not None and self._static_masks.is_cuda:
            self._packed_masks(self._static_masks, check=True)
        return self

    def _packed_masks(self, masks: Tensor, check: bool):
        """(bits, (Hm, Wm)) for `masks`, packed on first sight of this tensor object / version."""
        c = self._mask_cache
        if c is not None and c[0]() is masks and c[1] == masks._version:
            return c[2], c[3]
        # an expanded batch of identical masks (stride 0): pack one set
        src = masks[:1] if (masks.shape[0] > 1 and masks.stride(0) == 0) else masks
        bits = ops.pack_masks(src, check=check)
        hw = (int(masks.shape[3]), int(masks.shape[4]))
        self._mask_cache = (weakref.ref(masks), masks._version, bits, hw)
        return bits, hw

    @classmethod
    def from_reference(cls, module: nn.Module) -> "_SweepBase":
        """Build from a reference cv_builder instance, carrying over post_vol parameters and buffers."""
        k = module.post_vol.conv_layer.kernel_size[0]
        norm = {v: k_ for k_, v in NORM3D_TYPE.items()}.get(type(module.post_vol.norm_layer), "none")
        relu = {v: k_ for k_, v in RELU_TYPE.items()}.get(type(module.post_vol.activation), "none")
        new = cls(module.num_cams, module.feat_chs, k, norm, relu)
        new.post_vol.load_state_dict(module.post_vol.state_dict())
        new.train(module.training)
        return new.to(next(module.parameters()).device)

    def _run(self, feats: Tensor, grids, grid_masks, masks, want_valid: bool = False):
        if feats.shape[1] != self.num_cams:
            raise RuntimeError(f"expected {self.num_cams} cameras, feats has {feats.shape[1]}")
        if feats.dtype in (torch.float16, torch.bfloat16):
            # under torch.autocast / Lightning precision 16 the extractor emits half tensors; the sweep computes in
            # float32 like the reference's own sampler does under autocast (grid_sample is on the fp32 list)
            feats = feats.float()
        if self._rig_blob is not None:
            rig = self._rig
            D, Ho, Wo = rig.num_cands, rig.cv_camera.H, rig.cv_camera.W
            grids = grid_masks = None
        else:
            D, Ho, Wo = grids.shape[2:5]
        mask_h = mask_w = 0
        if self._aggregate != abi.AGG_STD:
            masks = None
        else:
            if self._static_masks is not None:
                masks = self._static_masks
            if masks is None:
                raise RuntimeError("the std builder needs masks")
            if masks.is_cuda and masks.dim() == 5 and masks.dtype in (torch.float32, torch.bool, torch.uint8):
                masks, (mask_h, mask_w) = self._packed_masks(masks, check=self.check_masks)
        order = table = None
        if ((self.balance_runs or self.keep_tap_table) and self._rig_blob is not None and not want_valid and
                feats.is_cuda and (masks is None or (self._static_masks is not None and masks.dtype == torch.int32))):
            order, table = self._octet_order(feats, masks, mask_h, mask_w, int(Ho), int(Wo), int(D))
        return ops.sweep_op(feats, grids, grid_masks, masks, self._rig_blob, self._aggregate, int(Ho), int(Wo), int(D),
                            want_valid, self.vol_channels_last, self.feats_bf16, mask_h, mask_w,
                            self.vol_bf16 and not want_valid, order, table)

    def _octet_order(self, feats: Tensor, bits: Optional[Tensor], mask_h: int, mask_w: int, Ho: int, Wo: int, D: int):
        """Cost-balanced assignment of the output image's octets to the sweep's CTAs (ops.balanced_octet_order) for
        the attached rig and the static masks: computed at the first call with a given shape (one device
        synchronisation, skipped while a CUDA graph is being captured) and kept, like the packed masks."""
        key = (tuple(feats.shape), tuple(feats.stride()), None if bits is None else (bits.data_ptr(), bits._version),
               mask_h, mask_w, Ho, Wo, D, self.vol_channels_last, self.feats_bf16, self.vol_bf16, feats.device,
               self.balance_runs, self.calibrate_runs, self.keep_tap_table)
        c = self._order_cache
        if c is not None and c[0] == key:
            return c[1], c[2]
        if torch.cuda.is_current_stream_capturing():
            return None, None
        need_run = (self.balance_runs and self.calibrate_runs) or self.keep_tap_table
        call = ops.SweepCall(feats, None, None, bits, self._rig_blob, self._aggregate, Ho, Wo, D,
                             describe_only=not need_run, channels_last=self.vol_channels_last,
                             feats_bf16=self.feats_bf16, mask_hw=(mask_h, mask_w) if mask_h > 0 else None,
                             vol_bf16=self.vol_bf16)
        order = None
        if self.balance_runs:
            if self.calibrate_runs:
                order, self.balance_report = ops.calibrated_octet_order(call)
            else:
                order = ops.balanced_octet_order(call)
        table = None
        if self.keep_tap_table:
            call.set_octet_order(order)
            table = ops.build_tap_table(call)  # follows the work assignment: built after it
        self._order_cache = (key, order, table)
        return order, table


class SphericalSweepStdMasked(_SweepBase):
    """Masked mean/variance sweep (reference spherical_sweep_avg.py:9)."""

    _aggregate = abi.AGG_STD

    def sweep(self, feats: Tensor, grids: Tensor, grid_masks: Tensor, masks: Tensor) -> Tensor:
        """feats [B,CAM,C,Hi,Wi], grids [B,CAM,X,Ho,Wo,2], grid_masks [B,CAM,X,Ho,Wo,1], masks [B,CAM,1,Hm,Wm]
        -> vol [B,C,X,Ho,Wo]   (reference :38-136)."""
        return self._run(feats, grids, grid_masks, masks)[0]

    def sweep_with_valid(self, feats, grids, grid_masks, masks):
        """Also returns the validity mask (n>1 cameras, reference :106-108) the reference never exposes."""
        vol, valid = self._run(feats, grids, grid_masks, masks, want_valid=True)
        return vol, valid.bool()

    def forward(self, feats: Tensor, grids: Tensor, grid_masks: Tensor, masks: Tensor) -> Tensor:
        vol = self.post_vol(self.sweep(feats, grids, grid_masks, masks))  # reference :162-167
        return vol.float() if self.vol_bf16 else vol


class SphericalSweep(_SweepBase):
    """Concat sweep (reference spherical_sweep.py:9); feat_chs = CAM * C of the extractor."""

    _aggregate = abi.AGG_CONCAT

    def sweep(self, feats: Tensor, grids: Tensor, masks: Optional[Tensor] = None) -> Tensor:
        """-> vol [B, CAM*C, X, Ho, Wo], channel = cam*C + c (reference :38-68); masks unused."""
        return self._run(feats, grids, None, None)[0]

    def forward(self, feats: Tensor, grids: Tensor, grid_masks: Tensor, masks: Tensor) -> Tensor:
        # grid_masks is not used; it is here for compatibility (reference :81).
        vol = self.post_vol(self.sweep(feats, grids, masks))
        return vol.float() if self.vol_bf16 else vol


class DistanceRegressorWithFixedCandidates(nn.Module):
    """Soft-argmin over fixed distance candidates (reference distance_regressor.py:7-80)."""

    def __init__(self, bf: float = 96, dist_cands: Sequence[float] = (0.5, 1, 1.5, 2, 5, 10, 20, 30, 50, 100),
                 interp_scale_factor: float = 0, pre_interp: bool = False):
        super().__init__()
        inv_dist_idx = bf / torch.tensor(list(dist_cands), dtype=torch.float32)
        self.register_buffer("inv_dist_idx", inv_dist_idx.view(1, -1, 1, 1), persistent=True)
        self.inv_dist_idx_min = float(torch.min(self.inv_dist_idx))
        self.inv_dist_idx_max = float(torch.max(self.inv_dist_idx))
        self.bf = bf
        self.interp_scale_factor = interp_scale_factor if interp_scale_factor > 0 else 0
        self.pre_interp = pre_interp
        # inference callers (torch_only.py:34, api/) drop norm_costs; set False to skip writing 6.5 MB/frame
        self.return_norm_costs = True

    @classmethod
    def from_reference(cls, module: nn.Module) -> "DistanceRegressorWithFixedCandidates":
        cands = (module.bf / module.inv_dist_idx.reshape(-1).double()).tolist()
        new = cls(module.bf, cands, module.interp_scale_factor, module.pre_interp)
        new.load_state_dict(module.state_dict())  # exact float32 buffer
        new.inv_dist_idx_min, new.inv_dist_idx_max = module.inv_dist_idx_min, module.inv_dist_idx_max
        return new.to(module.inv_dist_idx.device)

    def update_dist_cands(self, dist_cands: Sequence[float]) -> None:
        inv_dist_idx = self.bf / torch.tensor(list(dist_cands), dtype=torch.float32)
        self.inv_dist_idx[:] = inv_dist_idx.view(1, -1, 1, 1).to(dtype=self.inv_dist_idx.dtype,
                                                                device=self.inv_dist_idx.device)
        self.inv_dist_idx_min = float(torch.min(self.inv_dist_idx))
        self.inv_dist_idx_max = float(torch.max(self.inv_dist_idx))

    def _out_size(self, H: int, W: int):
        if self.pre_interp and self.interp_scale_factor > 0:
            s = self.interp_scale_factor
            # F.interpolate(scale_factor=s): floor(size * s); source-index scale = 1/s (recompute_scale_factor unset)
            return int(math.floor(float(H) * s)), int(math.floor(float(W) * s)), 1.0 / s
        return H, W, 1.0

    def regress(self, costs: Tensor, want_norm_costs: bool = True, want_argmax: bool = False):
        if costs.dtype in (torch.float16, torch.bfloat16):  # autocast: softmax runs in float32 there too
            costs = costs.float()
        H, W = costs.shape[-2:]
        Ho, Wo, scale = self._out_size(H, W)
        return ops.softargmin_op(costs, self.inv_dist_idx, Ho, Wo, float(scale), want_norm_costs, want_argmax)

    def forward(self, costs: Tensor):
        """costs [B,1,D,H,W] -> (inv_depths [B,1,sH,sW], norm_costs [B,D,sH,sW])  (reference :51-80)."""
        inv, nc, _ = self.regress(costs, want_norm_costs=self.return_norm_costs)
        return inv, (nc if self.return_norm_costs else None)


def patch_model(model: nn.Module, rig: Optional[Rig] = None, static_masks: Optional[Tensor] = None,
                return_norm_costs: bool = True) -> nn.Module:
    """Swap the reference's cv_builder / dist_regressor of a loaded model for the B200 modules, in place.

    Works on SphericalSweepStereoBase (dsta_mvs/model/mvs_model/torch_only.py:4-35) and on the Lightning wrappers that
    hold one as `.mvs_model` (spherical_sweep_stereo_2024.py) or hold the four sub-modules directly
    (spherical_sweep_stereo.py).  Checkpoints keep loading because parameter/buffer names are unchanged."""
    host = model.mvs_model if hasattr(model, "mvs_model") else model
    ref_b, ref_r = host.cv_builder, host.dist_regressor
    kind = type(ref_b).__name__
    if kind == "SphericalSweepStdMasked":
        new_b = SphericalSweepStdMasked.from_reference(ref_b)
    elif kind == "SphericalSweep":
        new_b = SphericalSweep.from_reference(ref_b)
    else:
        raise TypeError(f"cv_builder of type {kind} is not one of the two the reference ships")
    if rig is not None:
        new_b.attach_rig(rig)
    if static_masks is not None:
        new_b.set_static_masks(static_masks)
    new_r = DistanceRegressorWithFixedCandidates.from_reference(ref_r)
    new_r.return_norm_costs = return_norm_costs
    host.cv_builder, host.dist_regressor = new_b, new_r
    return model
