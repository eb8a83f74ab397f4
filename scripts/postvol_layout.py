"""post_vol (Conv3d 16->16 k3 + BatchNorm3d + LeakyReLU on cuDNN, spherical_sweep_avg.py:30-36,165) timed on the
volume the sweep kernel writes, in the reference's NCDHW layout and in channels_last_3d (MVSGI_FLAG_VOL_CHANNELS_LAST).
SURVEY.md §8 row f2: which hand-off layout makes sweep + post_vol fastest.  Run on the GPU box."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from mvs_gi_b200 import modules, synthetic

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
w = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "G16V"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rig = w.rig()
m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
m.set_static_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
feats = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev)
out = {"workload": w.name, "frames": B}
ref = None
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for cl in (False, True):
        m.set_vol_channels_last(cl)
        with torch.no_grad():
            vol = m.sweep(feats, None, None, None)
            for _ in range(3):
                y = m.post_vol(vol)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                y = m.post_vol(vol)
            ev[1].record()
            torch.cuda.synchronize()
            ms_post = ev[0].elapsed_time(ev[1]) / 10
            ev[0].record()
            for _ in range(10):
                y = m(feats, None, None, None)
            ev[1].record()
            torch.cuda.synchronize()
            ms_fwd = ev[0].elapsed_time(ev[1]) / 10
        if ref is None:
            ref = y.clone()
        err = float((y - ref).abs().max() / ref.abs().max())
        out[f"{'channels_last_3d' if cl else 'ncdhw'}{'_cudnn_benchmark' if bench else ''}"] = {
            "post_vol_ms": ms_post, "forward_ms": ms_fwd, "rel_err_vs_ncdhw": err,
            "out_channels_last": bool(y.is_contiguous(memory_format=torch.channels_last_3d))}
# the bfloat16 hand-off (MVSGI_FLAG_VOL_BF16): bf16 channels-last volume, post_vol in bfloat16 on the tensor cores
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    m.set_vol_bf16(True)
    with torch.no_grad():
        vol = m.sweep(feats, None, None, None)
        for _ in range(3):
            y = m.post_vol(vol)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            y = m.post_vol(vol)
        ev[1].record()
        torch.cuda.synchronize()
        ms_post = ev[0].elapsed_time(ev[1]) / 10
        ev[0].record()
        for _ in range(10):
            y = m(feats, None, None, None)
        ev[1].record()
        torch.cuda.synchronize()
        ms_fwd = ev[0].elapsed_time(ev[1]) / 10
        ev[0].record()
        for _ in range(10):
            vol = m.sweep(feats, None, None, None)
        ev[1].record()
        torch.cuda.synchronize()
        ms_sweep = ev[0].elapsed_time(ev[1]) / 10
    m.set_vol_bf16(False)
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    out[f"bf16_channels_last_3d{'_cudnn_benchmark' if bench else ''}"] = {
        "post_vol_ms": ms_post, "forward_ms": ms_fwd, "sweep_ms": ms_sweep, "rel_err_vs_fp32_ncdhw": err}
print(json.dumps(out))
