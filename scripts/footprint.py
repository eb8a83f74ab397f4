"""Offline analysis: unique feature pixels a tile of output pixels touches (all candidates, all cameras) at G16V."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from mvs_gi_b200 import synthetic
from oracle import oracle as orc

w = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "G16V"]
rs = w.rig().to_struct()
g, gm = orc.make_grids(rs)  # [CAM, D, Ho, Wo, 2]
CAM, D, Ho, Wo, _ = g.shape
Hf, Wf = w.feat_hw
x = ((g[..., 0] + 1) * Wf - 1) / 2
y = ((g[..., 1] + 1) * Hf - 1) / 2
x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64)
inside = (x >= -1) & (x < Wf) & (y >= -1) & (y < Hf)
print("step between consecutive candidates (feature px), per cam: median / 90% / max")
for c in range(CAM):
    dx = np.abs(np.diff(x[c], axis=0)); dx = np.minimum(dx, Wf - dx)
    for d in range(D - 1):
        pass
    print(c, [f"{np.median(dx[d]):.1f}/{np.percentile(dx[d],90):.1f}" for d in range(D - 1)])

def unique_px(tile_h, tile_w, dchunk):
    tot = 0; ntiles = 0
    for ty in range(0, Ho, tile_h):
        for tx in range(0, Wo, tile_w):
            for d0 in range(0, D, dchunk):
                for c in range(CAM):
                    sl = (c, slice(d0, d0 + dchunk), slice(ty, ty + tile_h), slice(tx, tx + tile_w))
                    m = inside[sl]
                    yy = y0[sl][m]; xx = x0[sl][m]
                    s = set()
                    keys = np.concatenate([(yy + dy) * 4096 + (xx + dx) for dy in (0, 1) for dx in (0, 1)])
                    tot += np.unique(keys).size
            ntiles += 1
    return tot * 64 / 1e6  # MB per frame (C=16 fp32)

for th, tw, dc in [(1, 88, 16), (1, 88, 4), (1, 88, 1), (1, 32, 16), (2, 44, 16), (4, 22, 16), (8, 11, 16), (8, 8, 16), (8, 8, 1), (4, 32, 16), (8, 40, 16), (16, 20, 16)]:
    print(f"tile {th}x{tw} cands/chunk {dc}: unique footprint {unique_px(th, tw, dc):8.1f} MB/frame  (feats {CAM*Hf*Wf*64/1e6:.1f} MB, vol 26.2 MB)")
