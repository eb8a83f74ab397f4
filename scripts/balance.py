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

# ---- the shipped assignment: octet-interleaved runs (run r owns octets r, r + runs, r + 2 runs, ...)
octs = (P + 7) // 8
oct_cost = per_px[: octs * 8].reshape(octs, 8).sum(1) + 0.3 * 16 * 8
for nruns in (291, 296):
    cost = np.array([oct_cost[r::nruns].sum() for r in range(nruns)])
    print(f"interleaved runs {nruns}: max/mean = {cost.max()/cost.mean():.3f}, min/mean = {cost.min()/cost.mean():.3f}, "
          f"std/mean = {cost.std()/cost.mean():.3f}")
    sm = cost[:148].copy()
    sm[: nruns - 148] += cost[148:]
    print(f"   per SM (CTA i and i + 148 together): max/mean = {sm.max()/sm.mean():.3f}, min/mean = {sm.min()/sm.mean():.3f}")
# greedy: octets sorted by cost, each to the currently lightest run with room (<= 11 octets)
order = np.argsort(-oct_cost)
nruns = 296
load = np.zeros(nruns); cnt = np.zeros(nruns, int)
import heapq
h = [(0.0, r) for r in range(nruns)]
heapq.heapify(h)
cap = (octs + nruns - 1) // nruns
for o in order:
    while True:
        l, r = heapq.heappop(h)
        if cnt[r] < cap:
            break
    load[r] += oct_cost[o]; cnt[r] += 1
    heapq.heappush(h, (load[r], r))
print(f"greedy LPT over {nruns} runs (<= {cap} octets each): max/mean = {load.max()/load.mean():.3f}, min/mean = {load.min()/load.mean():.3f}")
from mvs_gi_b200 import ops as _ops
_ord = _ops.lpt_assign((oct_cost * 10).astype(np.int64), 296, 11)
_load = np.array([oct_cost[o[o >= 0]].sum() for o in _ord])
print(f"round-wise LPT (ops.lpt_assign) over 296 runs: max/mean = {_load.max()/_load.mean():.3f}, min/mean = {_load.min()/_load.mean():.3f}")
