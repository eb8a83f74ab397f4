"""Diagnostic (GPU box): when does every CTA of the sweep start, finish its table build and finish
(MvsgiSweepArgs.cta_times_out) — which SMs / runs are the late ones?
usage: python scripts/cta_times.py WORKLOAD BATCH [0 = round-robin octets | 1 = cost-balanced | 2 = + measured correction]"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from mvs_gi_b200 import abi, ops, synthetic

name, B = sys.argv[1], int(sys.argv[2])
w = synthetic.WORKLOADS[name]
dev = torch.device("cuda:0")
rig_blob = ops.rig_to_blob(w.rig().to_struct())
masks = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
Hf, Wf = w.feat_hw
g = torch.Generator(device=dev).manual_seed(7)
sets = [torch.randn((B, w.num_cams, Hf, Wf, w.feat_chs), device=dev, generator=g).permute(0, 1, 4, 2, 3) for _ in range(3)]
calls = [ops.SweepCall(f, None, None, masks, rig_blob, abi.AGG_STD, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                       channels_last=True, mask_hw=w.mask_hw) for f in sets]
order = None
mode = sys.argv[3] if len(sys.argv) > 3 else "0"   # 0 = round-robin octets, 1 = cost-balanced, 2 = + measured correction
if mode == "1":
    order = ops.balanced_octet_order(calls[0])
elif mode == "2":
    order, rep = ops.calibrated_octet_order(calls[0])
    print("calibration:", json.dumps(rep))
bal = order is not None
if bal:
    for c in calls:
        c.set_octet_order(order)
if os.environ.get("TAPTABLE", "0") == "1":  # tap tables kept across calls: the timed launches copy them in
    table = ops.build_tap_table(calls[0])
    for c in calls[1:]:
        c.set_tap_table(table)
lib = abi.load_library()
for i in range(6):
    calls[i % 3].run()
torch.cuda.synchronize()
n = lib.mvsgi_sweep_grid(C.byref(calls[0].args))
bufs = [torch.zeros((n, 4), dtype=torch.int64, device=dev) for _ in calls]
for c, b_ in zip(calls, bufs):
    c.args.cta_times_out = b_.data_ptr()
# which CTA is the last one, launch after launch: the same run, the same SM, or a different one every time?
last = []
for i in range(8):
    calls[i % 3].run()
    torch.cuda.synchronize()
    tt = bufs[i % 3].cpu().numpy()
    d_ = (tt[:, 3] - tt[:, 1]) / 1e3
    top = np.argsort(-d_)[:3]
    last.append([(int(j), int(tt[j, 0]), round(float(d_[j]), 1)) for j in top] + [round(float(np.median(d_)), 1)])
print("slowest three (run, sm, us) + median per launch:")
for l in last:
    print("  ", l)
t = tt
sm, t0, tb, t1 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
start = t0.min()
dur = (t1 - t0) / 1e3
print(json.dumps({"workload": name, "batch": B, "balanced": bal, "ctas": int(len(t)), "kernel_us": float((t1.max() - start) / 1e3),
                  "start_skew_us": float((t0.max() - start) / 1e3), "build_us_mean": float(((tb - t0) / 1e3).mean()),
                  "cta_us_min": float(dur.min()), "cta_us_mean": float(dur.mean()), "cta_us_max": float(dur.max()),
                  "cta_us_p10_p50_p90": [float(np.percentile(dur, q)) for q in (10, 50, 90)]}))
# per SM: number of CTAs and the time its last CTA ends
ends = {}
for s, e, d in zip(sm, (t1 - start) / 1e3, dur):
    ends.setdefault(int(s), []).append((float(e), float(d)))
one = [v[0][1] for v in ends.values() if len(v) == 1]
two = [max(x[1] for x in v) for v in ends.values() if len(v) == 2]
print("SMs with 1 CTA:", len(one), "mean us", np.mean(one) if one else None, "| SMs with 2 CTAs:", len(two), "mean of slower CTA us",
      np.mean(two), "min", np.min(two), "max", np.max(two))
# does the duration follow the SM id (die / GPC) or the run index?
sm_dur = np.array([[s, max(x[1] for x in v)] for s, v in sorted(ends.items()) if len(v) == 2])
half = sm_dur[:, 0] < sm_dur[:, 0].max() / 2
print("slower-CTA us by SM id half: low", sm_dur[half, 1].mean(), "high", sm_dur[~half, 1].mean())
for lo in range(0, 160, 16):
    sel = (sm_dur[:, 0] >= lo) & (sm_dur[:, 0] < lo + 16)
    if sel.any():
        print(f"  SM {lo:3d}-{lo + 15:3d}: mean {sm_dur[sel, 1].mean():7.1f} us  max {sm_dur[sel, 1].max():7.1f}")
np.savez(f"gpurun_out/cta_times_{name}_{B}_{mode}.npz", t=t, order=(order.cpu().numpy() if order is not None else np.zeros(0)))
idx = np.arange(len(dur))
print("corr(duration, blockIdx) =", float(np.corrcoef(idx, dur)[0, 1]))
if order is not None or True:
    # cost of each run under the assignment in use
    runs, units, noct, nctas = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    if lib.mvsgi_sweep_balance_shape(C.byref(calls[0].args), C.byref(runs), C.byref(units), C.byref(noct), C.byref(nctas)) != 0:
        sys.exit(0)  # candidates swept in chunks: no per-run cost model
    cost = torch.empty((noct.value,), dtype=torch.int32, device=dev)
    lib.mvsgi_sweep_octet_cost(C.byref(calls[0].args), cost.data_ptr(), torch.cuda.current_stream().cuda_stream)
    c = cost.cpu().numpy().astype(np.int64)
    if order is not None:
        o = order.cpu().numpy()
        rc = np.array([c[r[r >= 0]].sum() for r in o])
    else:
        nr = len(t)
        rc = np.array([c[r::nr].sum() for r in range(nr)])
    m = min(len(rc), len(dur))
    print("run cost max/mean", rc.max() / rc.mean(), "corr(duration, run cost) =", float(np.corrcoef(rc[:m], dur[:m])[0, 1]))
