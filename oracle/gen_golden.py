#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):  python oracle/gen_golden.py
The reference is imported read-only; nothing from it is copied into the repo —
only the input/output tensors of its hot-path functions are stored, as small
fixtures that pin oracle/mvsgi_oracle.c (tests/test_oracle_golden.py) and, through
the oracle, the CUDA path (tests/test_gpu_parity.py also checks them directly).

Reference entry points exercised (paths relative to /root/reference):
  dsta_mvs/model/backports/backports.py:11            bilinear_grid_sample
  dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:38   SphericalSweepStdMasked.sweep
  dsta_mvs/model/cost_volume_builder/spherical_sweep.py:38       SphericalSweep.sweep
  dsta_mvs/model/distance_regressor/distance_regressor.py:51     DistanceRegressorWithFixedCandidates.forward
  dsta_mvs/support/dataset/torch_cuda_sweep.py:186,342,385       RayMaker / CameraModelGridMaker / transform_3D_points_torch
  dsta_mvs/mvs_utils/camera_models.py                            Equirectangular, DoubleSphere, LinearSphere
"""
import json
import os
import sys
from pathlib import Path

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from dsta_mvs.model.backports import bilinear_grid_sample
from dsta_mvs.model.cost_volume_builder import SphericalSweep, SphericalSweepStdMasked
from dsta_mvs.model.distance_regressor.distance_regressor import DistanceRegressorWithFixedCandidates
from dsta_mvs.mvs_utils.camera_models import DoubleSphere, Equirectangular, LinearSphere
from dsta_mvs.mvs_utils.ftensor import FTensor
from dsta_mvs.support.dataset.dist_label_table import FixedDistLabelTable
from dsta_mvs.support.dataset.torch_cuda_sweep import CameraModelGridMaker, RayMaker, transform_3D_points_torch

from mvs_gi_b200 import calibration as cal

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
torch.set_num_threads(4)

DS_CALIB = [  # eval_wrappers/dsta_sphere_stereo_cvpr2021/resources/calibration.json:39-83 (1216x1216)
    dict(fx=224.99704858314974, fy=222.61538110253663, cx=610.8194177583569, cy=612.733026847266,
         xi=-0.2798824735025879, alpha=0.5705641480250155),
    dict(fx=227.0079587821951, fy=224.89159992484438, cx=592.4624999094248, cy=625.9863162264933,
         xi=-0.2732279039025607, alpha=0.5723793857277628),
    dict(fx=451.02815256076249, fy=446.64278600601917, cx=615.4926934426732, cy=618.3938488717827,
         xi=0.4631993698034355, alpha=0.766170708696419),
    dict(fx=221.40629620252654, fy=219.26311511805566, cx=619.1665737685072, cy=603.8934432476885,
         xi=-0.29112734608076609, alpha=0.5676687440968526),
]


def ref_erp(H, W):
    return Equirectangular(dict(H=H, W=W), latitude_span=(-np.pi / 2, 0), open_span=False, in_to_tensor=True,
                           out_to_numpy=False)


def ref_grids(cv_model, cam_models, poses, dist_list):
    """create_grids_from_rays as the dataset does it (multi_view_camera_model_dataset.py:530-557)."""
    rays, _ = RayMaker(cv_model, "cv").make_rays(candidates=FixedDistLabelTable(dist_list, 96).dist())
    grids, masks = [], []
    for cam, T in zip(cam_models, poses):
        pose = FTensor(torch.as_tensor(T, dtype=torch.float32), f0="cv", f1="cif")
        inv = pose.inverse()
        pts = transform_3D_points_torch(inv.tensor().unsqueeze(0), rays.tensor().unsqueeze(0))
        g, m = CameraModelGridMaker(cam).make_grid(pts)
        grids.append(g.squeeze(0))
        masks.append(m.squeeze(0))
    return torch.stack(grids), torch.stack(masks)  # [CAM,D,Ho,Wo,2], [CAM,D,Ho,Wo,1]


def blob_masks(gen, B, CAM, Hm, Wm):
    """0/1 float masks with a disc-shaped valid region per camera (fisheye-footprint-like) and one erased
    rectangle per sample (mimics configs/augmentation/random_masking.yaml)."""
    yy, xx = np.mgrid[0:Hm, 0:Wm]
    m = np.zeros((B, CAM, 1, Hm, Wm), np.float32)
    for b in range(B):
        for c in range(CAM):
            cy, cx = Hm * (0.35 + 0.3 * gen.random()), Wm * (0.3 + 0.4 * gen.random())
            ry, rx = Hm * (0.45 + 0.2 * gen.random()), Wm * (0.3 + 0.15 * gen.random())
            disc = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            m[b, c, 0] = disc
            y0, x0 = int(gen.integers(0, Hm // 2)), int(gen.integers(0, Wm // 2))
            m[b, c, 0, y0:y0 + Hm // 6, x0:x0 + Wm // 6] = 0.0
    return m


def save(name, **arrays):
    path = OUT / f"{name}.npz"
    np.savez_compressed(path, **arrays)
    print(f"{path.name}: {path.stat().st_size / 1024:.0f} KiB  " +
          ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items() if hasattr(v, 'shape')))


def case_grid_sample(gen):
    im = gen.standard_normal((3, 5, 9, 14)).astype(np.float32)
    grid = gen.uniform(-1.35, 1.35, (3, 7, 11, 2)).astype(np.float32)
    # exact pixel centres / borders / far outside
    grid[0, 0, :4, 0] = [-1.0, 1.0, 0.0, (2 * 3 + 1) / 14 - 1]
    grid[0, 0, :4, 1] = [-1.0, 1.0, 0.0, (2 * 4 + 1) / 9 - 1]
    grid[1, 0, :3] = [[-5.0, 0.1], [0.2, 7.0], [-1.0 - 1e-6, 1.0 + 1e-6]]
    out = bilinear_grid_sample(torch.from_numpy(im), torch.from_numpy(grid), align_corners=False)
    save("grid_sample", im=im, grid=grid, out=out.numpy())


def synth_rig_case(num_cams, D_list, Ho, Wo, Hs, Ws, same_facing):
    poses = cal.fixture_poses(num_cams, same_facing=same_facing)
    g, m = ref_grids(ref_erp(Ho, Wo), [ref_erp(Hs, Ws) for _ in range(num_cams)], poses, D_list)
    return poses, g, m


def case_sweeps(gen):
    dist4 = [0.5, 0.892501608318949, 2.0467867239341255, 100.0]
    for name, CAM, facing in (("std3", 3, True), ("std4", 4, False)):
        B, C, Hf, Wf, Hm, Wm, Ho, Wo = 2, 8, 16, 64, 64, 256, 10, 40
        poses, g, gm = synth_rig_case(CAM, dist4, Ho, Wo, Hm, Wm, facing)
        grids = g.unsqueeze(0).repeat(B, 1, 1, 1, 1, 1).numpy().copy()
        # perturb frame 1 so the two frames differ and some samples leave the image
        grids[1] += gen.normal(0, 0.05, grids[1].shape).astype(np.float32)
        grids[1, :, :, :2] *= 1.2
        grid_masks = gm.unsqueeze(0).repeat(B, 1, 1, 1, 1, 1).numpy().copy()
        grid_masks[1] &= gen.random(grid_masks[1].shape) > 0.15
        feats = gen.standard_normal((B, CAM, C, Hf, Wf)).astype(np.float32)
        masks = blob_masks(gen, B, CAM, Hm, Wm)
        mod = SphericalSweepStdMasked(CAM, C, 3).eval()
        with torch.no_grad():
            vol = mod.sweep(torch.from_numpy(feats), torch.from_numpy(grids), torch.from_numpy(grid_masks),
                            torch.from_numpy(masks))
        frac = float((vol == 0).float().mean())
        print(f"  {name}: zero fraction {frac:.3f}")
        save(f"sweep_{name}", feats=feats, grids=grids, grid_masks=grid_masks, masks=masks, vol=vol.numpy(),
             poses=np.stack(poses), dist=np.array(dist4))
    # concat (E8/G8 style): 3 cams, masks unused
    B, CAM, C, Hf, Wf, Ho, Wo = 2, 3, 4, 16, 64, 10, 40
    dist3 = [0.5, 1.2684960362770972, 100.0]
    poses, g, gm = synth_rig_case(CAM, dist3, Ho, Wo, 64, 256, True)
    grids = g.unsqueeze(0).repeat(B, 1, 1, 1, 1, 1).numpy().copy()
    grids[1] = gen.uniform(-1.2, 1.2, grids[1].shape).astype(np.float32)
    feats = gen.standard_normal((B, CAM, C, Hf, Wf)).astype(np.float32)
    mod = SphericalSweep(CAM, CAM * C, 3).eval()
    with torch.no_grad():
        vol = mod.sweep(torch.from_numpy(feats), torch.from_numpy(grids), None)
    save("sweep_concat3", feats=feats, grids=grids, vol=vol.numpy())


def case_softargmin(gen):
    for name, D, dist, s, pre in (("sa_d16_s2", 16, cal.DIST_16CPD, 2, True), ("sa_d8_s0", 8, cal.DIST_8CPD, 0, False),
                                  ("sa_d8_s2_nopre", 8, cal.DIST_8L, 2, False), ("sa_d5_s3", 5, [0.5, 1, 2, 5, 100.0], 3, True)):
        costs = (gen.standard_normal((2, 1, D, 10, 40)) * 3).astype(np.float32)
        reg = DistanceRegressorWithFixedCandidates(96, dist, s, pre).eval()
        with torch.no_grad():
            inv, nc = reg(torch.from_numpy(costs))
        save(f"softargmin_{name}", costs=costs, inv_dist=inv.numpy(), norm_costs=nc.numpy(),
             argmax=nc.argmax(dim=1).numpy().astype(np.int32), inv_dist_idx=reg.inv_dist_idx.numpy().reshape(-1),
             dist=np.array(dist), scale=np.array(s), pre_interp=np.array(pre))


def case_geometry():
    Ho, Wo = 20, 80
    dist = [0.5, 0.7174058117693126, 1.637348559910646, 7.764432455310304, 100.0]
    meta = {}
    # (a) ERP surrogates 512x2048, 4-camera fixture rig with rotations
    poses = cal.fixture_poses(4, same_facing=False)
    g, m = ref_grids(ref_erp(Ho, Wo), [ref_erp(512, 2048)] * 4, poses, dist)
    save("geom_erp4", grids=g.numpy(), grid_masks=m.numpy(), poses=np.stack(poses), dist=np.array(dist))
    # (b) same-facing 3 cams (paper rig), ERP surrogates
    poses = cal.fixture_poses(3, same_facing=True)
    g, m = ref_grids(ref_erp(Ho, Wo), [ref_erp(512, 2048)] * 3, poses, dist)
    save("geom_erp3_same", grids=g.numpy(), grid_masks=m.numpy(), poses=np.stack(poses), dist=np.array(dist))
    # (c) raw DoubleSphere fisheyes (config104-style "without surrogate input")
    poses = cal.fixture_poses(4, same_facing=False)
    cams = [DoubleSphere(c["xi"], c["alpha"], c["fx"], c["fy"], c["cx"], c["cy"], 195, dict(H=1216, W=1216),
                         in_to_tensor=True, out_to_numpy=False) for c in DS_CALIB]
    g, m = ref_grids(ref_erp(Ho, Wo), cams, poses, dist)
    save("geom_ds4", grids=g.numpy(), grid_masks=m.numpy(), poses=np.stack(poses), dist=np.array(dist))
    meta["ds_calib"] = DS_CALIB
    # (d) LinearSphere 195 deg, 3 cams same-facing
    poses = cal.fixture_poses(3, same_facing=True)
    cams = [LinearSphere(195, dict(H=686, W=686), in_to_tensor=True, out_to_numpy=False) for _ in range(3)]
    g, m = ref_grids(ref_erp(Ho, Wo), cams, poses, dist)
    save("geom_ls3", grids=g.numpy(), grid_masks=m.numpy(), poses=np.stack(poses), dist=np.array(dist))
    (OUT / "geom_meta.json").write_text(json.dumps(meta, indent=1))


def case_backward():
    """gradients torch.autograd derives from the reference's forward code, on the inputs of the
    forward fixtures (which must exist already) and seeded upstream gradients.  Own RNG: does not disturb the others."""
    gen = np.random.default_rng(20261018)
    for name, kind in (("std3", "std"), ("std4", "std"), ("concat3", "concat")):
        z = np.load(OUT / f"sweep_{name}.npz")
        feats = torch.from_numpy(z["feats"]).requires_grad_(True)
        grids = torch.from_numpy(z["grids"])
        B, CAM, C = feats.shape[:3]
        if kind == "std":
            mod = SphericalSweepStdMasked(CAM, C, 3).eval()
            vol = mod.sweep(feats, grids, torch.from_numpy(z["grid_masks"]), torch.from_numpy(z["masks"]))
        else:
            mod = SphericalSweep(CAM, CAM * C, 3).eval()
            vol = mod.sweep(feats, grids, None)
        assert np.array_equal(vol.detach().numpy(), z["vol"])
        gv = gen.standard_normal(tuple(vol.shape)).astype(np.float32)
        vol.backward(torch.from_numpy(gv))
        save(f"sweep_{name}_bwd", grad_vol=gv, grad_feats=feats.grad.numpy())
    for name in ("sa_d16_s2", "sa_d8_s0", "sa_d8_s2_nopre", "sa_d5_s3"):
        z = np.load(OUT / f"softargmin_{name}.npz")
        costs = torch.from_numpy(z["costs"]).requires_grad_(True)
        reg = DistanceRegressorWithFixedCandidates(96, z["dist"].tolist(), float(z["scale"]), bool(z["pre_interp"])).eval()
        inv, nc = reg(costs)
        g_inv = gen.standard_normal(tuple(inv.shape)).astype(np.float32)
        g_nc = gen.standard_normal(tuple(nc.shape)).astype(np.float32)
        (g_both,) = torch.autograd.grad([inv, nc], costs, [torch.from_numpy(g_inv), torch.from_numpy(g_nc)],
                                        retain_graph=True)
        (g_only_inv,) = torch.autograd.grad(inv, costs, torch.from_numpy(g_inv))
        save(f"softargmin_{name}_bwd", g_inv=g_inv, g_nc=g_nc, grad_costs=g_both.numpy(),
             grad_costs_inv_only=g_only_inv.numpy())


def main():
    if "--only-backward" in sys.argv:
        case_backward()
        return
    gen = np.random.default_rng(20261017)
    case_grid_sample(gen)
    case_sweeps(gen)
    case_softargmin(gen)
    case_geometry()
    case_backward()


if __name__ == "__main__":
    main()
