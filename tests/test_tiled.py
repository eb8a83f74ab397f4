"""The tile-pipelined sweep (csrc/sweep_tiled.cuh, the default kernel) against the CTA-resident-table sweep
(csrc/sweep.cuh, MVSGI_FLAG_CTA_TABLE): same inputs, bit-identical volumes and validity masks.  The CTA-table kernel is
the one the golden-vector and oracle tests of round 1 pinned; test_gpu_parity.py now runs the tiled kernel through
the same goldens, and this file holds the two kernels against each other on shapes that exercise ragged tiles,
candidate splitting, 1-4 cameras, both aggregations, both volume layouts and both mask modes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvs_gi_b200 import abi

    abi.load_library()
    return torch.device("cuda:0")


def _run(dev, w, feats, masks, aggregate, channels_last, cta_table, want_valid=False, grids=None, grid_masks=None):
    from mvs_gi_b200 import abi, ops

    rig_blob = None if grids is not None else ops.rig_to_blob(w.rig().to_struct())
    call = ops.SweepCall(torch.from_numpy(feats).to(dev), grids, grid_masks,
                         None if masks is None else torch.from_numpy(masks).to(dev), rig_blob, aggregate,
                         w.cv_hw[0], w.cv_hw[1], len(w.dist_list), want_valid=want_valid,
                         channels_last=channels_last, cta_table=cta_table)
    call.run()
    torch.cuda.synchronize()
    return call.vol, call.valid


CASES = [
    # name, small() overrides, B, aggregate, per-frame masks
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 3, "std", True),
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 3, "std", False),
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3), 2, "std", True),   # ragged tiles
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=5), 1, "std", False),  # candidate split
    ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 2, "std", True),
    ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), 2, "std", True),
    ("G16V_DS", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=5), 2, "std", False),
    ("G8", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=8), 3, "concat", False),
    ("G8", dict(feat_hw=(17, 70), cv_hw=(9, 37), n_cands=3), 1, "concat", False),
]


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name,kw,B,agg,per_frame", CASES)
def test_tiled_equals_cta_table(dev, name, kw, B, agg, per_frame, channels_last):
    from mvs_gi_b200 import abi, synthetic

    w = synthetic.small(synthetic.WORKLOADS[name], **kw)
    feats = synthetic.make_feats(w, B, seed=5)
    masks = None
    if agg == "std":
        base = synthetic.fisheye_footprint_masks(w)
        masks = synthetic.erase_rectangles(base, B, seed=6) if per_frame else base[None]
    aggregate = abi.AGG_STD if agg == "std" else abi.AGG_CONCAT
    vol_t, valid_t = _run(dev, w, feats, masks, aggregate, channels_last, cta_table=False, want_valid=agg == "std")
    vol_c, valid_c = _run(dev, w, feats, masks, aggregate, channels_last, cta_table=True, want_valid=agg == "std")
    assert vol_t.shape == vol_c.shape
    assert torch.equal(valid_t, valid_c)
    assert torch.equal(vol_t, vol_c), f"max |diff| {float((vol_t - vol_c).abs().max())}"
    if agg == "std":
        assert 0.02 < float(valid_t.float().mean()) < 0.999


def test_tiled_supplied_shared_grids(dev):
    """Grids supplied by the caller with batch size 1 (shared by all frames) also take the tiled kernel."""
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=8)
    B = 2
    feats = synthetic.make_feats(w, B, seed=7)
    masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, seed=8)
    grids, gm = ops.project_only(w.rig().to_struct(), dev)
    grids, gm = grids[None], gm[None]
    out = []
    for cta in (False, True):
        out.append(_run(dev, w, feats, masks, abi.AGG_STD, True, cta, want_valid=True, grids=grids, grid_masks=gm))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    fused = _run(dev, w, feats, masks, abi.AGG_STD, True, False, want_valid=True)
    assert torch.equal(out[0][0], fused[0]), "supplied grids == the kernel's own grids must give the same volume"


def test_tiled_full_size_equals_cta_table(dev):
    """G16V at BASELINE.json's full size, 4 frames, per-frame masks: the two kernels agree bit for bit."""
    from mvs_gi_b200 import abi, synthetic

    w = synthetic.WORKLOADS["G16V"]
    B = 4
    feats = synthetic.make_feats(w, B, seed=9)
    masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, seed=10)
    a = _run(dev, w, feats, masks, abi.AGG_STD, True, False, want_valid=True)
    b = _run(dev, w, feats, masks, abi.AGG_STD, True, True, want_valid=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
