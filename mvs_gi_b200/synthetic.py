"""Seeded synthetic workloads of the reference's shapes: rigs, masks, features, costs.

No dataset, checkpoint or reference code is needed: the rig is the reference's frame-graph fixture
(dsta_mvs/mvs_utils/test_data/frame_graph.json), cameras are the 512x2048 half-sphere Equirectangular surrogates of
configs/dataset/surrogate_models_3cam.yaml (or the RTSS DoubleSphere fisheyes for the raw-fisheye mode), candidate
lists come from experiment_configs/sweep_hp_config{24,25,29,103}.yaml.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import calibration as cal

# eval_wrappers/dsta_sphere_stereo_cvpr2021/resources/calibration.json:39-83 (1216 x 1216, 195 deg)
DS_CALIB = [
    dict(fx=224.99704858314974, fy=222.61538110253663, cx=610.8194177583569, cy=612.733026847266,
         xi=-0.2798824735025879, alpha=0.5705641480250155),
    dict(fx=227.0079587821951, fy=224.89159992484438, cx=592.4624999094248, cy=625.9863162264933,
         xi=-0.2732279039025607, alpha=0.5723793857277628),
    dict(fx=451.02815256076249, fy=446.64278600601917, cx=615.4926934426732, cy=618.3938488717827,
         xi=0.4631993698034355, alpha=0.766170708696419),
    dict(fx=221.40629620252654, fy=219.26311511805566, cx=619.1665737685072, cy=603.8934432476885,
         xi=-0.29112734608076609, alpha=0.5676687440968526),
]


@dataclass
class Workload:
    """One named configuration of BASELINE.json / SURVEY.md §8d."""

    name: str
    aggregate: str  # "std" | "concat"
    num_cams: int
    feat_chs: int
    dist_list: List[float]
    feat_hw: tuple = (128, 512)
    mask_hw: tuple = (512, 2048)
    cv_hw: tuple = (80, 320)
    scale: int = 2  # dist_regressor.interp_scale_factor (configs/base_model.yaml:33)
    bf: float = 96.0
    same_facing: bool = True
    camera: str = "erp"  # "erp" surrogate | "ds" raw fisheye

    @property
    def num_cands(self) -> int:
        return len(self.dist_list)

    @property
    def out_chs(self) -> int:
        return self.feat_chs if self.aggregate == "std" else self.feat_chs * self.num_cams

    def rig(self) -> cal.Rig:
        poses = cal.fixture_poses(self.num_cams, same_facing=self.same_facing)
        if self.camera == "erp":
            cams = [cal.EquirectCamera(*self.mask_hw) for _ in range(self.num_cams)]
        else:
            cams = [cal.DoubleSphereCamera(c["xi"], c["alpha"], c["fx"], c["fy"], c["cx"], c["cy"], 195, 1216, 1216)
                    for c in DS_CALIB[:self.num_cams]]
        return cal.Rig(cal.EquirectCamera(*self.cv_hw), cams, poses, self.dist_list)

    # ---- algorithmic byte counts, per frame ----
    def vol_bytes(self) -> int:
        return self.out_chs * self.num_cands * self.cv_hw[0] * self.cv_hw[1] * 4

    def feats_bytes(self) -> int:
        return self.num_cams * self.feat_chs * self.feat_hw[0] * self.feat_hw[1] * 4

    def masks_bytes(self, shared: bool) -> float:
        if self.aggregate != "std":
            return 0
        per = self.num_cams * self.mask_hw[0] * self.mask_hw[1] * 4
        return 0 if shared else per  # shared masks are read once per launch, not per frame

    def grids_bytes(self) -> int:
        n = self.num_cams * self.num_cands * self.cv_hw[0] * self.cv_hw[1]
        return n * 8 + (n if self.aggregate == "std" else 0)

    def softargmin_bytes(self, with_norm_costs: bool) -> int:
        hc, wc = self.cv_hw
        s = max(self.scale, 1)
        b = self.num_cands * hc * wc * 4 + s * s * hc * wc * 4
        if with_norm_costs:
            b += self.num_cands * s * s * hc * wc * 4
        return b


WORKLOADS = {
    # C1: G16V = config29: std builder, 16 geometry-informed candidates, 16-ch features, 3 cameras
    "G16V": Workload("G16V", "std", 3, 16, cal.DIST_16CPD),
    # C2: E8 / G8 = config24 / config25: concat builder (3 x 16 -> 48 channels), 8 candidates
    "E8": Workload("E8", "concat", 3, 16, cal.DIST_8L),
    "G8": Workload("G8", "concat", 3, 16, cal.DIST_8CPD),
    # C3: G16VV = config103 (faithful: 16-ch std) and the literal BASELINE.json reading (48-ch std)
    "G16VV": Workload("G16VV", "std", 3, 16, cal.DIST_16CPD),
    "G16VV_C48": Workload("G16VV_C48", "std", 3, 48, cal.DIST_16CPD),
    # C4: 4-camera rig with surrogate inputs (fixture rotations)
    "4CAM": Workload("4CAM", "std", 4, 16, cal.DIST_16CPD, same_facing=False),
    # C5: high-D stress, 64 candidates at 2x ERP resolution
    "HIGHD": Workload("HIGHD", "std", 3, 16, cal.dist_inverse_spaced(64), feat_hw=(256, 1024), mask_hw=(1024, 4096),
                      cv_hw=(160, 640)),
    # raw-fisheye mode (config104-style, no surrogate): DoubleSphere projection in-kernel
    "G16V_DS": Workload("G16V_DS", "std", 3, 16, cal.DIST_16CPD, mask_hw=(1216, 1216), camera="ds",
                        same_facing=False),
}


def small(w: Workload, feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands: Optional[int] = None,
          feat_chs: Optional[int] = None) -> Workload:
    """A shrunken copy of a workload that the CPU oracle finishes in well under a second."""
    import copy

    s = copy.copy(w)
    s.feat_hw, s.cv_hw = feat_hw, cv_hw
    s.mask_hw = mask_hw if w.camera == "erp" else w.mask_hw
    if n_cands is not None:
        idx = np.linspace(0, len(w.dist_list) - 1, n_cands).round().astype(int)
        s.dist_list = [w.dist_list[i] for i in idx]
    if feat_chs is not None:
        s.feat_chs = feat_chs
    s.name = w.name + "_small"
    return s


def fisheye_footprint_masks(w: Workload, fov_degree: float = 195.0) -> np.ndarray:
    """[CAM, 1, Hm, Wm] float32 0/1: the footprint of a `fov_degree` fisheye (optical axis = surrogate +z, plus a
    per-camera yaw) inside each half-sphere ERP surrogate — how the reference's masks come about
    (multi_view_camera_model_dataset.py:424-438 resamples the fisheye mask into the surrogate)."""
    Hm, Wm = w.mask_hw
    if w.camera != "erp":
        # raw fisheye image: valid disc of the 195-degree image circle
        yy, xx = np.mgrid[0:Hm, 0:Wm].astype(np.float32)
        r = np.hypot(xx + 0.5 - Wm / 2, yy + 0.5 - Hm / 2)
        m = (r <= 0.49 * min(Hm, Wm)).astype(np.float32)
        return np.broadcast_to(m, (w.num_cams, 1, Hm, Wm)).copy()
    lon = (np.arange(Wm, dtype=np.float64) + 0.5) / Wm * 2 * math.pi - math.pi
    lat = (np.arange(Hm, dtype=np.float64) + 0.5) / Hm * (math.pi / 2) - math.pi / 2
    cl = np.cos(lat)[:, None]
    out = np.empty((w.num_cams, 1, Hm, Wm), np.float32)
    for c in range(w.num_cams):
        yaw = math.radians(12.0 * (c - (w.num_cams - 1) / 2))
        z = cl * np.cos(lon - yaw)[None, :]  # component along the fisheye axis
        out[c, 0] = (z >= math.cos(math.radians(fov_degree / 2))).astype(np.float32)
    return out


def erase_rectangles(masks: np.ndarray, batch: int, seed: int) -> np.ndarray:
    """Per-sample masks: replicate [CAM,1,H,W] to [B,CAM,1,H,W] and erase one random rectangle (5-20 % of the image)
    in one random camera per sample — what RandomSingleImageMasking does in training
    (dsta_mvs/support/augmentation/augmentation.py:39-59, configs/augmentation/random_masking.yaml)."""
    gen = np.random.default_rng(seed)
    out = np.repeat(masks[None], batch, 0).copy()
    _, CAM, _, H, W = out.shape
    for b in range(batch):
        cam = int(gen.integers(0, CAM))
        frac = gen.uniform(0.05, 0.2)
        rh = max(1, int(H * math.sqrt(frac)))
        rw = max(1, int(W * math.sqrt(frac)))
        y0, x0 = int(gen.integers(0, H - rh + 1)), int(gen.integers(0, W - rw + 1))
        out[b, cam, 0, y0:y0 + rh, x0:x0 + rw] = 0.0
    return out


def make_feats(w: Workload, batch: int, seed: int) -> np.ndarray:
    """White-noise N(0,1) features [B,CAM,C,Hf,Wf]: the worst case for interpolation sensitivity."""
    gen = np.random.default_rng(seed)
    return gen.standard_normal((batch, w.num_cams, w.feat_chs, *w.feat_hw), dtype=np.float32)


def make_costs(w: Workload, batch: int, seed: int) -> np.ndarray:
    gen = np.random.default_rng(seed + 7)
    return gen.standard_normal((batch, 1, w.num_cands, *w.cv_hw), dtype=np.float32)
