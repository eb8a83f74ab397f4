"""Host-side calibration: camera models + rig -> the POD struct the kernels take.

This is the part of the reference's grid precompute that stays on the host:
turning camera intrinsics, camera poses and the candidate list into float32
scalars *cast exactly as the reference casts them*, so the in-kernel projection
 starts from the same numbers.

Reference:
  dsta_mvs/mvs_utils/camera_models.py      Equirectangular :577-772, DoubleSphere :388-534, LinearSphere :1427-1548
  dsta_mvs/mvs_utils/ftensor/ftensor.py     FTensor.inverse :454-476
  dsta_mvs/support/dataset/dist_label_table.py  FixedDistLabelTable :104-122
  dsta_mvs/support/dataset/multi_view_camera_model_dataset.py  create_grids_from_rays :530-557
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import abi

_F = np.float32


def _f32(x) -> float:
    return float(_F(x))


@dataclass
class EquirectCamera:
    """camera_models.py:577-678.  Defaults are the surrogate cameras of
    configs/dataset/surrogate_models_3cam.yaml (lat span [-pi/2, 0])."""

    H: int
    W: int
    longitude_span: Sequence[float] = (-math.pi, math.pi)
    latitude_span: Sequence[float] = (-math.pi / 2, 0.0)
    open_span: bool = False

    model = abi.CAM_EQUIRECT

    def params(self) -> List[float]:
        lon0, lon1 = self.longitude_span
        span = lon1 - lon0  # python double, camera_models.py:640
        if self.open_span:
            span = (self.W - 1) / self.W * span  # :646
        lo0, lo1 = _F(lon0), _F(lon1)  # float32 tensors :651-652
        la0, la1 = _F(self.latitude_span[0]), _F(self.latitude_span[1])
        return [float(lo0), float(lo1), _f32(span), float(la0), float(_F(la1 - la0))]

    def cv_params(self):
        """(lon0, lon_span_pixel, lat0, lat_range) used by pixel_2_ray :690-708."""
        p = self.params()
        return p[0], p[2], p[3], p[4]


@dataclass
class DoubleSphereCamera:
    """camera_models.py:388-408."""

    xi: float
    alpha: float
    fx: float
    fy: float
    cx: float
    cy: float
    fov_degree: float
    H: int
    W: int

    model = abi.CAM_DOUBLE_SPHERE

    def params(self) -> List[float]:
        a = self.alpha
        w1 = a / (1 - a) if a <= 0.5 else (1 - a) / a  # :402-404
        w2 = (w1 + self.xi) / math.sqrt(2 * w1 * self.xi + self.xi ** 2.0 + 1)  # :406
        fov_rad = self.fov_degree / 180.0 * math.pi  # deg2rad :20-22
        return [_f32(self.xi), _f32(a), _f32(1 - a), _f32(self.fx), _f32(self.fy), _f32(self.cx), _f32(self.cy),
                _f32(-w2), _f32(fov_rad / 2.0)]


@dataclass
class LinearSphereCamera:
    """camera_models.py:1427-1441 (square images only)."""

    fov_degree: float
    H: int
    W: int

    model = abi.CAM_LINEAR_SPHERE

    def params(self) -> List[float]:
        assert self.H == self.W, "LinearSphere must be square (camera_models.py:1441)"
        return [_f32(self.fov_degree / 180.0 * math.pi)]


def invert_pose_f32(T_cv_cam: np.ndarray):
    """FTensor.inverse (ftensor.py:470-476) in float32: R_i = R^T, t_i = -R^T t."""
    T = np.asarray(T_cv_cam, dtype=_F)
    R_i = np.ascontiguousarray(T[:3, :3].T)
    t_i = (-R_i) @ T[:3, 3]
    return R_i.astype(_F), t_i.astype(_F)


@dataclass
class Rig:
    """cv camera + cameras + poses + candidates = everything create_grids_from_rays needs."""

    cv_camera: EquirectCamera
    cameras: List[object]
    poses_cv_cam: List[np.ndarray]  # 4x4, pose of each camera image frame in the cv frame
    dist_list: Sequence[float]
    _struct: object = field(default=None, repr=False)

    def __post_init__(self):
        if not 1 <= len(self.cameras) <= abi.MAX_CAMS:
            raise ValueError(f"1..{abi.MAX_CAMS} cameras supported, got {len(self.cameras)}")
        if len(self.cameras) != len(self.poses_cv_cam):
            raise ValueError("one pose per camera")
        if not 1 <= len(self.dist_list) <= abi.MAX_CANDS:
            raise ValueError(f"1..{abi.MAX_CANDS} candidates supported, got {len(self.dist_list)}")

    @property
    def num_cams(self) -> int:
        return len(self.cameras)

    @property
    def num_cands(self) -> int:
        return len(self.dist_list)

    def dist_f32(self) -> np.ndarray:
        """FixedDistLabelTable.dist()  dist_label_table.py:119."""
        return np.array(self.dist_list, dtype=_F)

    def to_struct(self) -> abi.MvsgiRig:
        if self._struct is not None:
            return self._struct
        rig = abi.MvsgiRig()
        rig.num_cams = self.num_cams
        rig.num_cands = self.num_cands
        rig.cv_width, rig.cv_height = self.cv_camera.W, self.cv_camera.H
        rig.cv_lon0, rig.cv_lon_span_pixel, rig.cv_lat0, rig.cv_lat_range = self.cv_camera.cv_params()
        for i, (cam, T) in enumerate(zip(self.cameras, self.poses_cv_cam)):
            c = rig.cams[i]
            c.model, c.width, c.height = cam.model, cam.W, cam.H
            R_i, t_i = invert_pose_f32(T)
            for k in range(9):
                c.rot[k] = float(R_i.reshape(-1)[k])
            for k in range(3):
                c.trans[k] = float(t_i[k])
            for k, v in enumerate(cam.params()):
                c.p[k] = v
        for i, d in enumerate(self.dist_f32()):
            rig.dist[i] = float(d)
        self._struct = rig
        return rig


# --------------------------------------------------------------------------
# Candidate lists shipped with the reference's experiment configs.
DIST_16CPD = [0.5, 0.5436655530390901, 0.5933598538089077, 0.6505899686755604, 0.7174058117693126,
              0.7966693499997578, 0.892501608318949, 1.0110614299826677, 1.161983802830304, 1.3612365930547528,
              1.637348559910646, 2.0467867239341255, 2.7192485658566654, 4.033339718346582, 7.764432455310304,
              100.0]  # experiment_configs/sweep_hp_config29.yaml:3 (G16V), config103 (G16VV)
DIST_8L = [0.5, 0.5828476269775187, 0.6986027944111776, 0.87173100871731, 1.1589403973509933, 1.728395061728395,
           3.3980582524271843, 100.0]  # experiment_configs/sweep_hp_config24.yaml:3 (E8)
DIST_8CPD = [0.5, 0.6010332486645973, 0.7386307369519187, 0.9400701573768901, 1.2684960362770972,
             1.9108621368429788, 3.7735705522309764, 100.0]  # experiment_configs/sweep_hp_config25.yaml:3 (G8)


def dist_inverse_spaced(n: int, near: float = 0.5, far: float = 100.0) -> List[float]:
    """n candidates equally spaced in inverse distance (the high-D stress list)."""
    inv = np.linspace(1.0 / near, 1.0 / far, n)
    return [float(1.0 / v) for v in inv]


def _quat_to_R(x, y, z, w) -> np.ndarray:
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def fixture_poses(num_cams: int, same_facing: bool = False, arm: float = 0.6000000238418579) -> List[np.ndarray]:
    """T_cv_cam for the rig of the reference's frame-graph fixture
    (dsta_mvs/mvs_utils/test_data/frame_graph.json): camera body frames on
    0.6 m arms at yaw 0/90/180/270 deg, cv frame = rig image frame,
    cbf->cif = T_body_image.  `same_facing` keeps the translations but drops the
    relative rotation (the paper's co-facing fisheyes, README.md:86)."""
    R_bi = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])  # T_body_image
    body = [
        ((arm, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0)),
        ((0.0, arm, 0.0), (0.0, 0.0, 0.7071067690849304, 0.7071067690849304)),
        ((-arm, 0.0, 0.0), (0.0, 0.0, 1.0, -4.371138828673793e-08)),
        ((0.0, -arm, 0.0), (0.0, 0.0, 0.7071067690849304, -0.7071067690849304)),
    ]
    poses = []
    for pos, q in body[:num_cams]:
        T_rb_cb = np.eye(4)
        T_rb_cb[:3, :3] = np.eye(3) if same_facing else _quat_to_R(*q)
        T_rb_cb[:3, 3] = pos
        T_bi = np.eye(4)
        T_bi[:3, :3] = R_bi
        # T_cv_cam = T_rif_rbf @ T_rbf_cbf @ T_cbf_cif
        poses.append(np.linalg.inv(T_bi) @ T_rb_cb @ T_bi)
    return poses
