"""Row f1 measurement (run on the GPU box): backward of the fused sweep and of the soft-argmin at the G16V shape,
CUDA-event times and algorithmic bytes against the HBM roofline.

    python scripts/backward_bench.py [frames]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from mvs_gi_b200 import abi, modules, ops, synthetic

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = synthetic.WORKLOADS["G16V"]
rig = w.rig()
blob = ops.rig_to_blob(rig.to_struct())
masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]
sets = []
for s in range(3):
    f = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, requires_grad=True)
    gv = torch.randn((B, w.out_chs, w.num_cands, *w.cv_hw), device=dev)
    c = torch.randn((B, 1, w.num_cands, *w.cv_hw), device=dev, requires_grad=True)
    sets.append((f, gv, c))
reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
Ho, Wo, scale = reg._out_size(*w.cv_hw)
g_inv = torch.randn((B, 1, Ho, Wo), device=dev)
g_nc = torch.randn((B, w.num_cands, Ho, Wo), device=dev)


def timed(fn, iters=15):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sweep_bwd(i):
    f, gv, _ = sets[i % 3]
    return ops.sweep_backward_op(gv, f.detach(), None, None, masks, blob, abi.AGG_STD, w.cv_hw[0], w.cv_hw[1],
                                 w.num_cands)


def soft_bwd(i):
    _, _, c = sets[i % 3]
    return ops.softargmin_backward_op(g_inv, g_nc, c.detach(), reg.inv_dist_idx, Ho, Wo, float(scale))


def fwd_bwd(i):  # through autograd, the way training_step calls it
    f, gv, _ = sets[i % 3]
    f.grad = None
    vol, _ = ops.sweep_op(f, None, None, masks, blob, abi.AGG_STD, w.cv_hw[0], w.cv_hw[1], w.num_cands, False)
    vol.backward(gv)


def nhwc(f):
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


sets_cl = [nhwc(f.detach()) for f, _, _ in sets]


def sweep_bwd_cl(i):  # channels-last features: the gradient is scattered straight into a channels-last grad_feats
    return ops.sweep_backward_op(sets[i % 3][1], sets_cl[i % 3], None, None, masks, blob, abi.AGG_STD, w.cv_hw[0],
                                 w.cv_hw[1], w.num_cands)


l0 = ops.launch_counter()
ms_sweep = timed(sweep_bwd)
launches = (ops.launch_counter() - l0) / 18
l0 = ops.launch_counter()
ms_sweep_cl = timed(sweep_bwd_cl)
launches_cl = (ops.launch_counter() - l0) / 18
g_a, g_b = sweep_bwd(0), sweep_bwd_cl(0)
rel_cl = float((g_a - g_b).abs().max() / g_a.abs().max())
ms_soft = timed(soft_bwd)
ms_fb = timed(fwd_bwd)
peak = 6544.0
try:
    peak = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
staged = w.num_cams * ((w.feat_chs + 15) // 16) * (w.feat_hw[0] + 3) * (w.feat_hw[1] + 2) * 64
# reads grad_vol + feats (+ staged copy written and read) ; staged gradient zeroed, accumulated, read ; grad_feats written
sweep_bytes = B * (w.vol_bytes() + 2 * w.feats_bytes() + 2 * staged + 3 * staged)
soft_bytes = B * (w.num_cands * w.cv_hw[0] * w.cv_hw[1] * 4 * 2 + (1 + w.num_cands) * Ho * Wo * 4)
print(json.dumps({
    "workload": "G16V", "frames": B, "sweep_backward_ms": ms_sweep, "sweep_backward_launches": launches,
    "sweep_backward_algorithmic_gbs": sweep_bytes / (ms_sweep * 1e-3) / 1e9,
    "sweep_backward_frac_of_peak": sweep_bytes / (ms_sweep * 1e-3) / 1e9 / peak,
    "sweep_backward_channels_last_ms": ms_sweep_cl, "sweep_backward_channels_last_launches": launches_cl,
    "sweep_backward_channels_last_frac_of_peak": B * (w.vol_bytes() + 3 * w.feats_bytes()) / (ms_sweep_cl * 1e-3) / 1e9 / peak,
    "channels_last_vs_nchw_rel_diff": rel_cl,
    "softargmin_backward_ms": ms_soft, "softargmin_backward_algorithmic_gbs": soft_bytes / (ms_soft * 1e-3) / 1e9,
    "softargmin_backward_frac_of_peak": soft_bytes / (ms_soft * 1e-3) / 1e9 / peak,
    "sweep_forward_plus_backward_autograd_ms": ms_fb, "peak_gbs": peak}))
