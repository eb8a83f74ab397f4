"""compute-sanitizer target: the kernels added late in round 2 on small shapes (balanced sweep, sweep backward staged /
in place, octet cost, nearest resampling).  The TMA kernels are left out: memcheck's parameter instrumentation breaks
__grid_constant__ tensor maps."""
import numpy as np
import torch

from mvs_gi_b200 import abi, ops, synthetic, resample
from mvs_gi_b200 import calibration as cal

dev = torch.device("cuda:0")


def nhwc(f):
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


for name, kw in [("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 77), n_cands=16)),
                 ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(24, 80), n_cands=16)),
                 ("G8", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=8))]:
    w = synthetic.small(synthetic.WORKLOADS[name], **kw)
    std = w.aggregate == "std"
    agg = abi.AGG_STD if std else abi.AGG_CONCAT
    blob = ops.rig_to_blob(w.rig().to_struct())
    masks_f = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]
    bits = ops.pack_masks(masks_f) if std else None
    for B in (1, 5):
        f = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev)
        for feats in (f, nhwc(f)):
            call = ops.SweepCall(feats, None, None, bits, blob, agg, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                                 channels_last=True, mask_hw=tuple(w.mask_hw) if std else None)
            call.run()
            order = ops.balanced_octet_order(call)
            if order is not None:
                call.set_octet_order(order)
                call.run()
                sm, us = ops.cta_times(call, launches=1)
            gv = torch.randn((B, w.out_chs, len(w.dist_list), *w.cv_hw), device=dev)
            g = ops.sweep_backward_op(gv, feats, None, None, masks_f if std else None, blob, agg, w.cv_hw[0], w.cv_hw[1],
                                      len(w.dist_list))
            torch.cuda.synchronize()
            assert torch.isfinite(g).all()
raw = cal.DoubleSphereCamera(-0.28, 0.57, 28.0, 27.8, 76.0, 76.5, 195, 152, 152)
s = resample.CameraModelRotation(raw, cal.EquirectCamera(40, 160), np.eye(3))
img = torch.rand((2, 3, 152, 152), device=dev)
for mode in ("linear", "nearest"):
    out, valid = s(img, interpolation=mode)
torch.cuda.synchronize()
print("sanitize target done")
