"""compute-sanitizer target: the kept tap table (BUILD / USE launches of the sweep kernel: cp.async.bulk + mbarrier), the
item queue (more items than CTAs) and the rotated candidate chunks, on small shapes; every USE volume is compared with
the plain launch bit for bit."""
import torch

from mvs_gi_b200 import abi, ops, synthetic

dev = torch.device("cuda:0")


def nhwc(f):
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


for name, B, kw in [("G16V", 5, dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 77), n_cands=16)),
                    ("HIGHD", 2, dict(feat_hw=(64, 256), mask_hw=(256, 1024), cv_hw=(48, 160), n_cands=64)),  # 552 items: queue
                    ("4CAM", 3, dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(24, 80), n_cands=16)),
                    ("G8", 1, dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=8))]:
    w = synthetic.small(synthetic.WORKLOADS[name], **kw)
    std = w.aggregate == "std"
    agg = abi.AGG_STD if std else abi.AGG_CONCAT
    blob = ops.rig_to_blob(w.rig().to_struct())
    bits = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]) if std else None
    f = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev)
    for feats in (f, nhwc(f)):
        mk = lambda: ops.SweepCall(feats, None, None, bits, blob, agg, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                                   channels_last=True, mask_hw=tuple(w.mask_hw) if std else None)
        plain, kept = mk(), mk()
        plain.run()
        table = ops.build_tap_table(kept)
        assert table is not None
        for _ in range(2):
            kept.vol.zero_()
            kept.run()
        torch.cuda.synchronize()
        assert torch.equal(plain.vol.view(torch.int32), kept.vol.view(torch.int32)), name
        items = kept.lib.mvsgi_sweep_grid(__import__("ctypes").byref(kept.args))
        print(name, B, "ctas", items, "table bytes", table.numel(), flush=True)
print("sanitize target done")
