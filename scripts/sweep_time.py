"""Times the sweep kernel alone (static packed masks, channels-last features tapped in place, channels-last volume)
with the library MVSGI_B200_LIB selects, and prints a checksum of the volume bits for A/B bit-identity checks.
usage: sweep_time.py WORKLOAD BATCH [steps]"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from mvs_gi_b200 import abi, ops, synthetic

name, B = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
w = synthetic.WORKLOADS[name]
dev = torch.device("cuda:0")
rig_blob = ops.rig_to_blob(w.rig().to_struct())
masks = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
Hf, Wf = w.feat_hw
Ho, Wo = w.cv_hw
g = torch.Generator(device=dev).manual_seed(7)
nsets = max(3, int(200e6 // (B * w.num_cams * Hf * Wf * w.feat_chs * 4)) + 1)
sets = [torch.randn((B, w.num_cams, Hf, Wf, w.feat_chs), device=dev, generator=g).permute(0, 1, 4, 2, 3)
        for _ in range(min(nsets, 8))]
if os.environ.get("SAMEFRAME"):  # every frame of the batch is the same frame (stride 0): no HBM feature reads after the first
    sets = [f[:1].expand(B, -1, -1, -1, -1) for f in sets]
agg = abi.AGG_STD if w.aggregate == "std" else abi.AGG_CONCAT
calls = [ops.SweepCall(f, None, None, masks if agg == abi.AGG_STD else None, rig_blob, agg, Ho, Wo, len(w.dist_list),
                       channels_last=True, mask_hw=w.mask_hw) for f in sets]
balanced = False
if os.environ.get("BALANCE", "1") != "0":  # cost-balanced runs (ops.balanced_octet_order), computed once
    if os.environ.get("BALANCE") == "2":  # + measured correction
        order, rep = ops.calibrated_octet_order(calls[0], rounds=int(os.environ.get("CAL_ROUNDS", "3")),
                                                launches=int(os.environ.get("CAL_LAUNCHES", "3")),
                                                threshold=float(os.environ.get("CAL_THRESHOLD", "0.015")))
        print("calibration:", json.dumps(rep), flush=True)
    else:
        order = ops.balanced_octet_order(calls[0])
    if order is not None:
        balanced = True
        for c in calls:
            c.set_octet_order(order)
kept = False
if os.environ.get("TAPTABLE", "0") == "1":  # tap tables kept across calls (ops.build_tap_table): one BUILD launch here
    table = ops.build_tap_table(calls[0])
    if table is not None:
        kept = True
        for c in calls[1:]:
            c.set_tap_table(table)
n = len(calls)
for i in range(2 * n):
    calls[i % n].run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    calls[i % n].run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
calls[0].run()
torch.cuda.synchronize()
vol = calls[0].vol
chk = int(vol.contiguous().view(torch.int32).to(torch.int64).sum().item())
cout = vol.shape[1]
alg = B * (cout * len(w.dist_list) * Ho * Wo * 4 + w.num_cams * Hf * Wf * w.feat_chs * 4)
print(json.dumps({"lib": os.path.basename(os.environ.get("MVSGI_B200_LIB", "default")), "workload": name, "batch": B, "balanced": balanced, "tap_table_kept": kept,
                  "sweep_ms": round(ms, 4), "frac": round(alg / (ms * 1e-3) / 6537.6e9, 4), "checksum": chk}), flush=True)
