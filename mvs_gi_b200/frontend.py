"""Calibration front-end: the reference's JSON / YAML calibration files -> `calibration.Rig`.

The reference turns `manifest.json` (camera models + samplers), `frame_graph.json` (rig extrinsics) and the
`map_cam_frame_*` / `surrogate_models_*` / `cam_model_cv` configs into per-camera sampling grids on the CPU at
dataset construction (api/inference_class.py:37-45 builds a dummy dataloader just for that, and :54-79 documents that
the grids differ between machines).  Here the same files become the ~760-byte `MvsgiRig` the fused kernel evaluates
in-kernel; `ops.project_only` dumps the grids for callers that still want them.

Reference (paths relative to the checkout):
  dsta_mvs/mvs_utils/frame_io.py:10-137                 frame_graph.json grammar (frames / typical_poses / transforms)
  dsta_mvs/mvs_utils/ftensor/frame_graph.py:100-167     edges hold T_parent_child, inverse on the reverse edge;
                                                        query = shortest path, float32 product along the chain
  dsta_mvs/mvs_utils/ftensor/ftensor.py:454-476         FTensor.inverse (R^T, -R^T t)
  dsta_mvs/support/dataset/grid_maker_from_manifest.py:24-38   read_manifest
  dsta_mvs/mvs_utils/camera_models.py:32-48,381-408,577-678,1427-1441   make_object and the camera constructors
  dsta_mvs/support/dataset/multi_view_camera_model_dataset.py:519-557   pose of camera k = query(cv frame, cif_k)
  configs/dataset/{surrogate_models_3cam,cam_model_cv,map_cam_frame_*}.yaml

Third-party arithmetic restated here: pyquaternion 0.9.x `Quaternion(w, x, y, z).rotation_matrix` (frame_io.py:10-12;
the package is not installed in this image) — the unit-quaternion rotation matrix, evaluated in float64 after
normalisation and cast to float32 like `rot_mat_from_quat` does.
"""
from __future__ import annotations

import json
import math
from collections import deque
from pathlib import Path
from typing import Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np

from . import calibration as cal

_F = np.float32
PI_LATITUDE_SPAN = (-math.pi / 2, 0.0)  # the "sweep@pi_latitude_span" variable of the reference's configs


def _quat_rotation_matrix(x: float, y: float, z: float, w: float) -> np.ndarray:
    """pyquaternion's rotation_matrix for the normalised quaternion (float64), cast to float32 (frame_io.py:10-12)."""
    q = np.array([w, x, y, z], dtype=np.float64)
    n = np.linalg.norm(q)
    if n == 0:
        raise ValueError("zero quaternion")
    w, x, y, z = q / n
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)
    return R.astype(_F)


def _parse_orientation(o: Mapping) -> np.ndarray:
    """frame_io.py:18-41."""
    t = o["type"]
    if t == "rotation_matrix":
        return np.array(o["data"], dtype=_F).reshape(3, 3)
    if t == "quaternion":
        d = o["data"]
        return _quat_rotation_matrix(d["x"], d["y"], d["z"], d["w"])
    raise ValueError(f"Unsupported type of orientation. type = {t}. ")


def _parse_pose(p: Mapping) -> Tuple[np.ndarray, np.ndarray]:
    return np.array(p["position"], dtype=_F), _parse_orientation(p["orientation"])


def _inverse_f32(T: np.ndarray) -> np.ndarray:
    """FTensor.inverse (ftensor.py:470-476) in float32."""
    inv = np.eye(4, dtype=_F)
    R = np.ascontiguousarray(T[:3, :3].T)
    inv[:3, :3] = R
    inv[:3, 3] = (-R) @ T[:3, 3]
    return inv


class FrameGraph:
    """Reference frames and the transforms between them (frame_graph.py:60-167).  An edge (f0, f1) holds T_f0_f1, the
    pose of frame f1 in frame f0, as a float32 4x4; the reverse edge holds its FTensor inverse."""

    def __init__(self):
        self.frames: Dict[str, str] = {}
        self.edges: Dict[str, Dict[str, np.ndarray]] = {}

    def add_frame(self, name: str, comment: str = "") -> None:
        self.frames[name] = comment
        self.edges.setdefault(name, {})

    def add_or_update_pose_edge(self, f0: str, f1: str, T: np.ndarray) -> None:
        for f in (f0, f1):
            if f not in self.frames:
                self.add_frame(f)
        T = np.asarray(T, dtype=_F)
        self.edges[f0][f1] = T
        self.edges[f1][f0] = _inverse_f32(T)

    def shortest_path(self, f0: str, f1: str) -> List[str]:
        if f0 not in self.frames or f1 not in self.frames:
            raise KeyError(f"{f0 if f0 not in self.frames else f1} does not exist in the frame graph. ")
        prev = {f0: None}
        dq = deque([f0])
        while dq:
            u = dq.popleft()
            if u == f1:
                break
            for v in self.edges[u]:
                if v not in prev:
                    prev[v] = u
                    dq.append(v)
        if f1 not in prev:
            raise KeyError(f"no path between {f0} and {f1}")
        path = [f1]
        while path[-1] != f0:
            path.append(prev[path[-1]])
        return path[::-1]

    def query_transform(self, f0: str, f1: str) -> np.ndarray:
        """T_f0_f1: float32 product of the edge transforms along the shortest path (frame_graph.py:115-167)."""
        if f0 == f1:
            if f0 not in self.frames:
                raise KeyError(f"{f0} does not exist in the frame graph. ")
            return np.eye(4, dtype=_F)
        path = self.shortest_path(f0, f1)
        tf = np.eye(4, dtype=_F)
        for p, c in zip(path[:-1], path[1:]):
            tf = (tf @ self.edges[p][c]).astype(_F)
        return tf


def parse_frame_graph(obj: Mapping) -> FrameGraph:
    """frame_io.py:43-137 on an already-loaded JSON object."""
    G = FrameGraph()
    for entry in obj["frames"]:
        G.add_frame(entry["name"], entry.get("comment", ""))
    typical = {k: _parse_pose(v) for k, v in obj.get("typical_poses", {}).items()}
    for entry in obj["transforms"]:
        pose = entry["pose"]
        if pose["type"] == "create":
            position, orientation = _parse_pose(pose)
        elif pose["type"] == "reference":
            position, orientation = typical[pose["key"]]
        else:
            raise ValueError(f"Unsupported pose type. type = {pose['type']}. ")
        T = np.eye(4, dtype=_F)
        T[:3, :3] = orientation
        T[:3, 3] = position
        G.add_or_update_pose_edge(entry["f0"], entry["f1"], T)
    return G


def read_frame_graph(path) -> FrameGraph:
    with open(path, "r") as fp:
        return parse_frame_graph(json.load(fp))


def read_manifest(path) -> Tuple[dict, dict]:
    """grid_maker_from_manifest.py:24-38: (camera_models, {mvs_cam_key: sampler spec of the main camera model})."""
    with open(path, "r") as fp:
        obj = json.load(fp)
    samplers = {}
    for spec in obj["samplers"]:
        if spec["mvs_main_cam_model_for_cam"]:
            samplers[spec["mvs_cam_key"]] = spec["sampler"]
    return obj["camera_models"], samplers


def _shape(spec) -> Tuple[int, int]:
    ss = spec["shape_struct"]
    if isinstance(ss, Mapping):
        return int(ss["H"]), int(ss["W"])
    return int(ss.H), int(ss.W)


def _span(v, default):
    if v is None:
        return default
    if isinstance(v, str):
        if v.endswith("pi_latitude_span"):  # "sweep@pi_latitude_span", resolved by the reference's config helper
            return PI_LATITUDE_SPAN
        raise ValueError(f"unresolved config variable {v!r}")
    return tuple(float(x) for x in v)


def camera_from_spec(spec: Mapping):
    """camera_models.make_object (camera_models.py:32-48) for the three models the sweep can project into."""
    t = spec["type"]
    H, W = _shape(spec)
    if t == "Equirectangular":  # :580-586 (defaults: lon [-pi, pi], lat [-pi/2, pi/2])
        return cal.EquirectCamera(H, W, _span(spec.get("longitude_span"), (-math.pi, math.pi)),
                                  _span(spec.get("latitude_span"), (-math.pi / 2, math.pi / 2)),
                                  bool(spec.get("open_span", False)))
    if t == "DoubleSphere":  # :383-391
        return cal.DoubleSphereCamera(spec["xi"], spec["alpha"], spec["fx"], spec["fy"], spec["cx"], spec["cy"],
                                      spec["fov_degree"], H, W)
    if t == "LinearSphere":  # :1429-1441
        return cal.LinearSphereCamera(spec["fov_degree"], H, W)
    raise ValueError(f"camera model {t!r} is not one the sweep kernel projects into "
                     f"(Equirectangular, DoubleSphere, LinearSphere)")


def rig_from_calibration(frame_graph: Union[FrameGraph, str], map_cam_frame: Mapping[str, str],
                         camera_specs: Mapping[str, Mapping], cv_spec: Mapping, dist_list: Sequence[float],
                         extra_transforms: Optional[Sequence[Tuple[str, str, np.ndarray]]] = None) -> cal.Rig:
    """Build the kernel's rig from the reference's calibration inputs.

    frame_graph     : FrameGraph or path of a frame_graph.json (mvs_utils/frame_io.py grammar)
    map_cam_frame   : configs/dataset/map_cam_frame_*.yaml — {"cam0": frame, ..., "cv": frame} ("rig" is ignored)
    camera_specs    : the camera model each camK is swept with: configs/dataset/surrogate_models_*.yaml entries, or the
                      manifest's own camera_models for raw-fisheye sweeping (camera_models.make_object dicts)
    cv_spec         : configs/dataset/cam_model_cv.yaml
    dist_list       : the candidate distances (experiment_configs/*.yaml `dist_list`)
    extra_transforms: (f0, f1, T_f0_f1) edges added first — the surrogate frames the dataset creates at run time
                      (image_sampler/camera_model_sampler.py: R_raw_fisheye) when they are not in the JSON file.
    Pose of camera k = query_transform(cv frame, frame of camK) (multi_view_camera_model_dataset.py:519-527)."""
    G = read_frame_graph(frame_graph) if isinstance(frame_graph, (str, bytes)) or hasattr(frame_graph, "__fspath__") \
        else frame_graph
    for f0, f1, T in (extra_transforms or ()):
        G.add_or_update_pose_edge(f0, f1, T)
    cam_keys = sorted((k for k in map_cam_frame if k.startswith("cam")), key=lambda k: int(k[3:]))
    if not cam_keys:
        raise ValueError("map_cam_frame names no camK entries")
    cv_frame = map_cam_frame["cv"]
    cv_cam = camera_from_spec(cv_spec)
    if not isinstance(cv_cam, cal.EquirectCamera):
        raise ValueError("the cost-volume camera must be Equirectangular (configs/dataset/cam_model_cv.yaml)")
    cams, poses = [], []
    for k in cam_keys:
        cams.append(camera_from_spec(camera_specs[k]))
        poses.append(G.query_transform(cv_frame, map_cam_frame[k]).astype(np.float64))
    return cal.Rig(cv_cam, cams, poses, list(dist_list))


def rig_from_manifest(manifest_path, frame_graph, map_cam_frame: Mapping[str, str], cv_spec: Mapping,
                      dist_list: Sequence[float], surrogate_specs: Optional[Mapping[str, Mapping]] = None,
                      extra_transforms=None) -> cal.Rig:
    """SurrogateCameraModelGridMaker semantics (grid_maker_from_manifest.py:141-192): camera K is swept with its
    surrogate model when `surrogate_specs` names it, else with the manifest's camera model for that camera."""
    camera_models, samplers = read_manifest(manifest_path)
    specs = {}
    for cam_key, sampler in samplers.items():
        if not cam_key.startswith("cam"):
            continue
        if surrogate_specs and cam_key in surrogate_specs:
            specs[cam_key] = surrogate_specs[cam_key]
        else:
            specs[cam_key] = camera_models[sampler["cam_model_key"]]
    return rig_from_calibration(frame_graph, {k: v for k, v in map_cam_frame.items() if k == "cv" or k in specs},
                                specs, cv_spec, dist_list, extra_transforms)


# ---- masks.json -> surrogate validity masks ----------------------------------------------------------------------------
def read_mask(fn, valid_value: int = 255) -> np.ndarray:
    """Gray image == valid_value -> bool [H, W]  (mvs_utils/image_io/image_read.py:22-28, cv2.IMREAD_GRAYSCALE)."""
    try:
        import cv2

        img = cv2.imread(str(fn), cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise FileNotFoundError(f"Failed to read {fn}. ")
    except ImportError:  # PIL's 'L' conversion equals OpenCV's for single-channel PNGs, which is what masks are
        from PIL import Image

        img = np.asarray(Image.open(str(fn)).convert("L"))
    return img == valid_value


def read_masks(dataset_path, mask_path="masks.json"):
    """multi_view_camera_model_dataset.py:223-277: {"masks": [{"raw_camera": "cam0", "mask": "valid_mask_0.png"}, ...],
    "masks_augmented": [...] (optional)} -> (masks, masks_augmented): dicts raw_camera -> float32 0/1 array [H, W]
    (masks_augmented is None when the file has no such section)."""
    root = Path(dataset_path)
    with open(root / mask_path, "r") as fp:
        mask_dict = json.load(fp)

    def section(name):
        return {d["raw_camera"]: read_mask(root / d["mask"]).astype(np.float32) for d in mask_dict[name]}

    return section("masks"), (section("masks_augmented") if "masks_augmented" in mask_dict else None)


def sample_masks(masks: Mapping[str, np.ndarray], samplers: Mapping[str, object], device="cuda"):
    """The surrogate masks the sweep consumes: every raw fisheye mask through its camera's resampler
    (multi_view_camera_model_dataset.py:424-438): `mask * 255` sampled with invalid value 0, then `> 0 -> 1`.
    `samplers[cam]` is a mvs_gi_b200.resample.CameraModelRotation (the f4 kernel) or None for cameras that are swept
    raw (the reference's NoOpSampler: the mask is passed through).  Returns dict cam -> float32 CUDA tensor [1, Hm, Wm]."""
    import torch

    out = {}
    for cam, mask in masks.items():
        m = torch.as_tensor(np.asarray(mask, dtype=np.float32), device=device)[None]  # ToTensor: [1, H, W]
        sampler = samplers.get(cam)
        if sampler is not None:
            m, _ = sampler(m * 255, invalid_pixel_value=0)
            m = (m > 0).to(torch.float32).squeeze(0)
        out[cam] = m
    return out


def stack_masks(masks: Mapping[str, object], cam_keys: Optional[Sequence[str]] = None):
    """dict cam -> [1, Hm, Wm]  ->  [1, CAM, 1, Hm, Wm], cameras in camK order: what `set_static_masks` / the `masks`
    argument of the sweep modules take (the DataLoader stacks them the same way, :721-731)."""
    import torch

    keys = list(cam_keys) if cam_keys is not None else sorted(masks, key=lambda k: int(k[3:]))
    return torch.stack([masks[k] for k in keys], dim=0)[None]
