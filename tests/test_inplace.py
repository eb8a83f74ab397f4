"""Channels-last features are tapped where they are (no staging pass): same bits as the zero-framed staged copy.

The sweep kernel (csrc/sweep.cuh) reads staged features — zero-framed, channel-group-major, written by
stage_feats_kernel — when the caller hands over NCHW (what cuDNN produces by default), and the caller's own tensor when
it is channels-last with whole 16-channel groups (north_star (a)); the tap-table entries then carry clamped offsets and
edge flags (make_entry_direct) instead of relying on a zero frame.  The staged path is the one the golden-vector and
oracle tests pin (tests/test_gpu_parity.py); this file holds the in-place path against it on shapes that exercise
image edges, ragged sizes, 1-4 cameras, both aggregations, both volume layouts and both mask modes."""
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


def _nhwc(f):
    """The same [B, CAM, C, Hf, Wf] tensor with channels-last strides (what an NHWC extractor hands over)."""
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


def _run(dev, w, feats, masks, aggregate, channels_last, feats_nhwc=False, staged_feats=False, want_valid=False):
    from mvs_gi_b200 import ops

    f = torch.from_numpy(feats).to(dev)
    if feats_nhwc:
        f = _nhwc(f)
    call = ops.SweepCall(f, None, None, None if masks is None else torch.from_numpy(masks).to(dev),
                         ops.rig_to_blob(w.rig().to_struct()), aggregate, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                         want_valid=want_valid, channels_last=channels_last, staged_feats=staged_feats)
    call.run()
    torch.cuda.synchronize()
    return call.vol, call.valid, call


CASES = [
    # name, small() overrides, B, aggregate, per-frame masks
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 3, "std", True),
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=5), 2, "std", False),
    ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 2, "std", True),
    ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), 2, "std", True),
    ("G16V_DS", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=5), 2, "std", False),
    ("G8", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=8), 3, "concat", False),
    ("G8", dict(feat_hw=(5, 9), cv_hw=(9, 37), n_cands=3), 1, "concat", False),  # nearly every tap touches an edge
    ("G16V", dict(feat_hw=(2, 3), mask_hw=(8, 12), cv_hw=(6, 24), n_cands=4), 2, "std", False),  # both edges at once
]


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name,kw,B,agg,per_frame", CASES)
def test_channels_last_features_tapped_in_place(dev, name, kw, B, agg, per_frame, channels_last):
    from mvs_gi_b200 import abi, synthetic

    w = synthetic.small(synthetic.WORKLOADS[name], **kw)
    feats = synthetic.make_feats(w, B, seed=15)
    masks = None
    if agg == "std":
        base = synthetic.fisheye_footprint_masks(w)
        masks = synthetic.erase_rectangles(base, B, seed=16) if per_frame else base[None]
    aggregate = abi.AGG_STD if agg == "std" else abi.AGG_CONCAT
    wv = agg == "std"
    ref = _run(dev, w, feats, masks, aggregate, channels_last, want_valid=wv)                      # NCHW, staged
    direct = _run(dev, w, feats, masks, aggregate, channels_last, feats_nhwc=True, want_valid=wv)  # in place
    staged = _run(dev, w, feats, masks, aggregate, channels_last, feats_nhwc=True, staged_feats=True, want_valid=wv)
    assert direct[2].launch_count() == ref[2].launch_count() - 1, "no staging launch for channels-last features"
    if wv:
        assert torch.equal(direct[1], ref[1]) and torch.equal(staged[1], ref[1])
    assert torch.equal(staged[0], ref[0])
    assert torch.equal(direct[0], ref[0]), f"max |diff| {float((direct[0] - ref[0]).abs().max())}"
    if agg == "std":
        assert 0.02 < float(ref[1].float().mean()) < 0.999


def test_in_place_call_needs_no_feature_workspace(dev):
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=4)
    f = torch.from_numpy(synthetic.make_feats(w, 2, seed=1)).to(dev)
    m = torch.from_numpy(synthetic.fisheye_footprint_masks(w)[None]).to(dev)
    blob = ops.rig_to_blob(w.rig().to_struct())
    nchw = ops.SweepCall(f, None, None, m, blob, abi.AGG_STD, *w.cv_hw, 4)
    nhwc = ops.SweepCall(_nhwc(f), None, None, m, blob, abi.AGG_STD, *w.cv_hw, 4)
    staged_bytes = 2 * 3 * (32 + 3) * (128 + 2) * 16 * 4
    assert nchw.workspace.numel() - nhwc.workspace.numel() >= staged_bytes


def test_full_size_in_place_equals_staged(dev):
    """G16V at BASELINE.json's full size, 4 frames, per-frame masks."""
    from mvs_gi_b200 import abi, synthetic

    w = synthetic.WORKLOADS["G16V"]
    B = 4
    feats = synthetic.make_feats(w, B, seed=9)
    masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, seed=10)
    a = _run(dev, w, feats, masks, abi.AGG_STD, True, want_valid=True)
    b = _run(dev, w, feats, masks, abi.AGG_STD, True, feats_nhwc=True, want_valid=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])


def test_channels_last_slices_and_odd_channel_counts_fall_back_to_staging(dev):
    """Layouts the in-place path does not take (a channel slice of a wider tensor keeps the pixel pitch, 12 channels
    are not a whole group) still give the staged result."""
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=4,
                        feat_chs=12)
    feats = synthetic.make_feats(w, 2, seed=3)
    masks = synthetic.fisheye_footprint_masks(w)[None]
    ref = _run(dev, w, feats, masks, abi.AGG_STD, False)
    odd = _run(dev, w, feats, masks, abi.AGG_STD, False, feats_nhwc=True)
    assert odd[2].launch_count() == ref[2].launch_count()  # staged
    assert torch.equal(odd[0], ref[0])
    # 16 channels sliced out of a 32-channel channels-last tensor: pixel pitch 32, in place
    w16 = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=4)
    wide = torch.randn((2, 3, 16, 64, 32), device=dev)
    sl = wide[..., 8:24].permute(0, 1, 4, 2, 3)  # [B, CAM, 16, Hf, Wf], channel stride 1, base offset 32 bytes
    m = torch.from_numpy(masks).to(dev)
    blob = ops.rig_to_blob(w16.rig().to_struct())
    c1 = ops.SweepCall(sl, None, None, m, blob, abi.AGG_STD, *w16.cv_hw, 4)
    c1.run()
    c2 = ops.SweepCall(sl.contiguous(), None, None, m, blob, abi.AGG_STD, *w16.cv_hw, 4)
    c2.run()
    torch.cuda.synchronize()
    assert torch.equal(c1.vol, c2.vol)


# Several frames per pass (sweep.cuh, FB = 2 / 3 / 4): chosen by the library for static masks + channels-last features
# tapped in place + channels-last volume without a validity output.  Asking for the validity output selects the
# one-frame-per-pass instantiation, so the two volumes must be bit-identical; batches that are not a multiple of FB end
# on a partial pass.
FB_CASES = [
    # name, small() overrides, B (-> FB per sweep_frames_per_pass), aggregate
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 2, "std"),      # FB 2
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 5, "std"),      # FB 2, tail
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=5), 18, "std"),         # FB 4, tail 2
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=5), 17, "std"),         # FB 4, tail 1
    ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), 7, "std"),      # FB 3, tail 1
    ("4CAM", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), 3, "std"),         # FB 2, tail
    ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), 16, "std"),   # 3 groups: FB 2
    ("G16V_DS", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=5), 16, "std"),                       # FB 4
    ("G8", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=8), 19, "concat"),                         # FB 4, tail 3
    ("G8", dict(feat_hw=(5, 9), cv_hw=(9, 37), n_cands=3), 3, "concat"),                              # FB 2, edges
]


@pytest.mark.parametrize("vol_bf16", [False, True])
@pytest.mark.parametrize("name,kw,B,agg", FB_CASES)
def test_frames_per_pass_instantiations_are_bit_identical(dev, name, kw, B, agg, vol_bf16):
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS[name], **kw)
    f = _nhwc(torch.from_numpy(synthetic.make_feats(w, B, seed=21)).to(dev))
    aggregate = abi.AGG_STD if agg == "std" else abi.AGG_CONCAT
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)[None]).to(dev) if agg == "std" else None
    blob = ops.rig_to_blob(w.rig().to_struct())

    def run(**kwargs):
        call = ops.SweepCall(f, None, None, masks, blob, aggregate, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                             channels_last=True, **kwargs)
        call.vol.fill_(-7.0)  # a frame the partial pass must not touch twice nor skip
        call.run()
        torch.cuda.synchronize()
        return call.vol

    if vol_bf16:
        fast = run(vol_bf16=True)
        ref = run(vol_bf16=True, staged_feats=True)   # staged features: one frame per pass
    else:
        fast = run()
        ref = run(want_valid=True) if agg == "std" else run(staged_feats=True)
    assert torch.equal(fast, ref), f"max |diff| {float((fast.float() - ref.float()).abs().max())}"
    assert float(fast.float().abs().max()) > 0
