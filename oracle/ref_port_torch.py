"""PyTorch-eager restatement of the reference path, for the TIMED CPU BASELINE only.  TEST INFRASTRUCTURE.

The reference is Python/PyTorch and cannot travel to the GPU box (/root/reference does not exist there), so
`bench.py --impl reference` and bench.py's `cpu_baseline` time this port instead (kind = "port").  It deliberately
keeps the reference's *formulation* — a Python loop over candidates, a zero-pad + four index-gathers sampler, masked
mean / variance with torch.where — because that formulation (not the arithmetic) is what the reference's CPU cost
comes from: ~45 tensor ops per sampler call, each materialising a [N, C, Ho*Wo] intermediate.

Checked against tests/golden/ (bit-exact for the sweeps) in tests/test_oracle_golden.py.  Never imported by
mvs_gi_b200.

Follows (paths relative to the reference checkout):
  dsta_mvs/model/backports/backports.py:11-86
  dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:38-136
  dsta_mvs/model/cost_volume_builder/spherical_sweep.py:38-68
  dsta_mvs/model/distance_regressor/distance_regressor.py:51-80
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _clip(idx: torch.Tensor, hi: int) -> torch.Tensor:
    lo_c = torch.tensor(0, device=idx.device)
    hi_c = torch.tensor(hi, device=idx.device)
    idx = torch.where(idx < 0, lo_c, idx)
    return torch.where(idx > hi, hi_c, idx)


def sample_bilinear_zero_pad(image: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """image [N,C,H,W], grid [N,Hg,Wg,2] in [-1,1] (align_corners=False) -> [N,C,Hg,Wg]   (backports.py:11-86)."""
    n, c, h, w = image.shape
    hg, wg = grid.shape[1:3]
    px = (((grid[..., 0] + 1) * w - 1) / 2).reshape(n, -1)  # :41-42
    py = (((grid[..., 1] + 1) * h - 1) / 2).reshape(n, -1)
    left = torch.floor(px).long()
    top = torch.floor(py).long()
    right, bottom = left + 1, top + 1
    w_tl = ((right - px) * (bottom - py)).unsqueeze(1)  # :52-55
    w_bl = ((right - px) * (py - top)).unsqueeze(1)
    w_tr = ((px - left) * (bottom - py)).unsqueeze(1)
    w_br = ((px - left) * (py - top)).unsqueeze(1)
    padded = F.pad(image, pad=[1, 1, 1, 1], mode="constant", value=0)  # :58 (whole image, every call)
    pw, ph = w + 2, h + 2
    left, right = _clip(left + 1, pw - 1), _clip(right + 1, pw - 1)  # :62-72
    top, bottom = _clip(top + 1, ph - 1), _clip(bottom + 1, ph - 1)
    flat = padded.view(n, c, -1)

    def take(xi, yi):
        lin = (xi + yi * pw).unsqueeze(1).expand(-1, c, -1)  # :76-84
        return torch.gather(flat, 2, lin)

    out = take(left, top) * w_tl + take(left, bottom) * w_bl + take(right, top) * w_tr + take(right, bottom) * w_br
    return out.reshape(n, c, hg, wg)


def sweep_std(feats, grids, grid_masks, masks):
    """spherical_sweep_avg.py:38-136 -> [B, C, D, Ho, Wo]."""
    b, cams, d_n, ho, wo = grids.shape[:5]
    c = feats.shape[2]
    f_flat, m_flat, g_flat = feats.flatten(0, 1), masks.flatten(0, 1), grids.flatten(0, 1)
    slices = []
    for d in range(d_n):
        g_ok = grid_masks[:, :, d, :, :, 0].view(b, cams, 1, ho, wo)
        warped = sample_bilinear_zero_pad(f_flat, g_flat[:, d]).view(b, cams, c, ho, wo)
        seen = (sample_bilinear_zero_pad(m_flat, g_flat[:, d]) > 0.0).view(b, cams, 1, ho, wo)
        seen = torch.logical_and(seen, g_ok)
        n_seen = torch.sum(seen.to(torch.float), dim=1, keepdim=True)
        usable = n_seen > 1.0
        denom = torch.where(usable, n_seen, torch.ones(usable.shape, dtype=n_seen.dtype, device=usable.device))
        mean = torch.sum(warped * seen.to(torch.float), dim=1, keepdim=True) / denom
        warped = torch.where(seen, warped, mean)
        var = torch.sum((warped - mean) ** 2, dim=1, keepdim=True) / denom
        var = torch.where(usable.expand_as(var), var, torch.zeros(var.shape, dtype=var.dtype, device=var.device))
        slices.append(var[:, 0].unsqueeze(2))
    return torch.cat(slices, dim=2)


def sweep_concat(feats, grids):
    """spherical_sweep.py:38-68 -> [B, CAM*C, D, Ho, Wo]."""
    b, cams, d_n, ho, wo = grids.shape[:5]
    c = feats.shape[2]
    f_flat, g_flat = feats.flatten(0, 1), grids.flatten(0, 1)
    slices = []
    for d in range(d_n):
        warped = sample_bilinear_zero_pad(f_flat, g_flat[:, d]).view(b, cams, c, ho, wo)
        slices.append(warped.flatten(1, 2).unsqueeze(2))
    return torch.cat(slices, dim=2)


def soft_argmin(costs, inv_dist_idx, interp_scale_factor=0, pre_interp=False):
    """distance_regressor.py:51-80 -> (inv_dist [B,1,H',W'], norm_costs [B,D,H',W'])."""
    vol = costs[:, 0]
    if pre_interp and interp_scale_factor > 0:
        vol = F.interpolate(vol, size=None, scale_factor=interp_scale_factor, mode="bilinear")
    prob = F.softmax(vol, 1)
    inv = torch.sum(prob * inv_dist_idx.view(1, -1, 1, 1).expand_as(prob), dim=1, keepdim=True)
    return inv, prob
