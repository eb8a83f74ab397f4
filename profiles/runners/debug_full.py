import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from mvs_gi_b200 import modules, ops, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "HIGHD"
dev = torch.device("cuda:0")
w = synthetic.WORKLOADS[name]
rig = w.rig()
gen = torch.Generator(device=dev).manual_seed(1234)
feats = torch.randn((1, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)[None]).to(dev)
m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
with torch.no_grad():
    vol, valid = m.sweep_with_valid(feats, None, None, masks)
nz = vol != 0
bad = nz != valid[:, None].expand_as(vol)
print("mismatches", int(bad.sum()), "of", bad.numel())
idx = bad.nonzero()
print(idx[:20].tolist())
print("valid but zero:", int((valid[:, None].expand_as(vol) & ~nz).sum()), " nonzero but invalid:", int((nz & ~valid[:, None].expand_as(vol)).sum()))
for i in idx[:5].tolist():
    print(i, vol[tuple(i)].item(), vol[i[0], :, i[2], i[3], i[4]].tolist())
