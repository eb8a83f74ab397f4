"""Offline: per-run work (valid camera taps) for the G16V bench inputs -> load imbalance across CTAs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from mvs_gi_b200 import synthetic
from oracle import oracle as orc

w = synthetic.WORKLOADS["G16V"]
rs = w.rig().to_struct()
g, gm = orc.make_grids(rs)
masks = synthetic.fisheye_footprint_masks(w)
CAM, D, Ho, Wo, _ = g.shape
# validity per (cam, d, h, w) = grid valid & bilinear(mask) > 0, via the oracle sampler on the mask image
valid = np.zeros((CAM, D, Ho, Wo), bool)
for c in range(CAM):
    for d in range(D):
        s = orc.bilinear_grid_sample(masks[c][None], g[c, d][None])[0, 0]
        valid[c, d] = (s > 0) & gm[c, d]
n = valid.sum(0)  # [D, Ho, Wo]
taps = np.where(n > 1, n, 0)  # cameras actually sampled per (d, px)
print("valid voxels %.1f%%, mean cams per voxel %.2f" % (100 * (n > 1).mean(), taps.mean()))
P = Ho * Wo
per_px = taps.reshape(D, P).sum(0)  # cam-taps per pixel over all candidates (max 48)
for nruns in (291, 296, 592, 1184):
    octs = (P + 7) // 8
    cost = []
    for r in range(nruns):
        p0 = 8 * (octs * r // nruns); p1 = min(P, 8 * (octs * (r + 1) // nruns))
        cost.append(per_px[p0:p1].sum() + 0.3 * 16 * (p1 - p0))  # + store/overhead term
    cost = np.array(cost)
    print(f"runs {nruns}: max/mean = {cost.max()/cost.mean():.2f}, min/mean = {cost.min()/cost.mean():.2f}")
    if nruns in (291,):
        # pairing runs r and r+148 on the same SM?
        pass
# row profile
rowcost = per_px.reshape(Ho, Wo).sum(1)
print("row cost (first 10, mid, last):", rowcost[:5], rowcost[38:42], rowcost[-5:])
colcost = per_px.reshape(Ho, Wo).sum(0)
print("col cost by 32-px block:", colcost.reshape(10, 32).sum(1))
