"""Row f4 measurement (run on the GPU box): fisheye -> surrogate-ERP resampling, 1216 x 1216 RTSS fisheyes ->
512 x 2048 surrogates, against the HBM roofline, with the reference's formulation (a stored grid + F.grid_sample per
camera, api/inference_pytorch.py:61-74) timed on the same GPU next to it."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.nn.functional as F

from mvs_gi_b200 import calibration as cal
from mvs_gi_b200 import ops, resample, synthetic

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24  # images per launch (8 frames x 3 cameras)
c = synthetic.DS_CALIB[0]
raw = cal.DoubleSphereCamera(c["xi"], c["alpha"], c["fx"], c["fy"], c["cx"], c["cy"], 195, 1216, 1216)
tgt = cal.EquirectCamera(512, 2048)
R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], dtype=np.float32)
s = resample.CameraModelRotation(raw, tgt, R)
imgs = [torch.rand((N, 3, 1216, 1216), device=dev) for _ in range(3)]  # 3 x 425 MB: each set larger than L2


def timed(fn, iters=20):
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


ms = timed(lambda i: s(imgs[i % 3]))
# camera buffers as they arrive: uint8 HWC, permuted view, scaled on load (fuses inference_pytorch.py:55-57)
u8 = [(im * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous() for im in imgs]
ms_u8 = timed(lambda i: s(u8[i % 3].permute(0, 3, 1, 2)))
ms_u8_ref = timed(lambda i: s(u8[i % 3].permute(0, 3, 1, 2).float().div_(255.0)), iters=10)
out, valid = s(imgs[0])
grids, gm = ops.project_only(s._struct, dev)  # [1,1,Ht,Wt,2]: the grid the reference would have stored
grid = grids[0, 0][None]
grid = torch.where(gm[0, 0], grid, torch.full_like(grid, -1.0))


def ref_step(i):
    outs = []
    for n in range(N):  # the reference loops over cameras / frames, batch 1 each
        o = F.grid_sample(imgs[i % 3][n:n + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        o[..., ~valid] = 0.0
        outs.append(o)
    return torch.cat(outs, 0)


ms_ref = timed(ref_step, iters=5)
ref = ref_step(0)
err = float((ref - out).abs().max())
out_bytes = N * 3 * 512 * 2048 * 4
in_bytes = N * 3 * 1216 * 1216 * 4
peak = 6544.0
try:
    peak = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
gbs = (out_bytes + in_bytes) / (ms * 1e-3) / 1e9
print(json.dumps({"images_per_launch": N, "kernel_ms": ms, "images_per_s": N / (ms * 1e-3),
                  "algorithmic_bytes": out_bytes + in_bytes, "achieved_gbs": gbs, "peak_gbs": peak,
                  "frac": gbs / peak, "valid_fraction": float(valid.float().mean()),
                  "uint8_hwc_kernel_ms": ms_u8, "uint8_hwc_convert_then_kernel_ms": ms_u8_ref,
                  "uint8_hwc_achieved_gbs": (out_bytes + in_bytes // 4) / (ms_u8 * 1e-3) / 1e9,
                  "torch_grid_sample_loop_ms": ms_ref, "speedup_vs_torch_loop": ms_ref / ms,
                  "max_abs_diff_vs_torch_grid_sample": err}))
