"""Debug helper (GPU box): compare the fused sweep with the CPU oracle on a shrunken workload, print mismatch stats."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from mvs_gi_b200 import modules, ops, synthetic
from oracle import oracle as orc

name = sys.argv[1] if len(sys.argv) > 1 else "HIGHD"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
w = synthetic.small(synthetic.WORKLOADS[name], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80))
rig = w.rig()
feats = synthetic.make_feats(w, B, 3)
masks = synthetic.fisheye_footprint_masks(w)[None]
m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
with torch.no_grad():
    vol, valid = m.sweep_with_valid(torch.from_numpy(feats).to(dev), None, None, torch.from_numpy(masks).to(dev))
    grids, gm = ops.project_only(rig.to_struct(), dev)
ref, ref_valid = orc.sweep_std(feats, grids.cpu().numpy()[None], gm.cpu().numpy()[None], masks, return_valid=True)
vol, valid = vol.cpu().numpy(), valid.cpu().numpy()
print("valid mismatch per d:", (valid != ref_valid).sum(axis=(0, 2, 3)))
print("zero-pattern mismatch per d:", ((vol == 0) != (ref == 0)).sum(axis=(0, 1, 3, 4)))
print("max err per d:", np.abs(vol - ref).max(axis=(0, 1, 3, 4)))
