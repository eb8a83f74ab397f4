"""Tap tables kept across calls (MvsgiSweepArgs.tap_table, ops.build_tap_table).

The sweep kernel's shared-memory records — output offset, validity bits, per camera the tap offset and the two bilinear
fractions — depend on the rig, the masks, the shapes and the work assignment, not on the features: they are to this
kernel what the grids the reference's dataset precomputes once per rig (multi_view_camera_model_dataset.py:519-557) are
to `SphericalSweepStdMasked.forward` (spherical_sweep_avg.py:138-167).  A BUILD launch stores them, USE launches copy
them in (cp.async.bulk + mbarrier) instead of evaluating the geometry.  The volume must not change by a bit."""
import ctypes as C

import numpy as np
import pytest
import torch


def test_sweep_args_carry_the_tap_table_fields():
    from mvs_gi_b200 import abi

    names = [f[0] for f in abi.MvsgiSweepArgs._fields_]
    assert names[-4:] == ["tap_table", "tap_table_bytes", "tap_table_mode", "reserved2"]
    assert (abi.TAP_TABLE_OFF, abi.TAP_TABLE_BUILD, abi.TAP_TABLE_USE) == (0, 1, 2)
    assert "mvsgi_sweep_tap_table_bytes" in abi.EXPORTS


def test_tap_table_size_is_planned_without_a_gpu():
    """mvsgi_sweep_tap_table_bytes only plans: G16V keeps 44-byte records for 25 600 pixels x 16 candidates in 296
    slots (18.3 MB), per-frame masks keep nothing, and a launch that asks for a table it cannot keep is refused before
    anything is enqueued."""
    from mvs_gi_b200 import abi, synthetic

    lib = abi.load_library()
    w = synthetic.WORKLOADS["G16V"]
    rig = w.rig().to_struct()
    a = abi.MvsgiSweepArgs()
    a.struct_size = C.sizeof(a)
    a.aggregate, a.grid_mode, a.flags = abi.AGG_STD, abi.GRID_PROJECT, abi.FLAG_VOL_CHANNELS_LAST
    a.batch, a.num_cams, a.feat_chs, a.num_cands = 32, w.num_cams, w.feat_chs, len(w.dist_list)
    a.feat_h, a.feat_w = w.feat_hw
    a.out_h, a.out_w = w.cv_hw
    a.mask_h, a.mask_w = w.mask_hw
    Hf, Wf, Cc = w.feat_hw[0], w.feat_hw[1], w.feat_chs
    a.feats = 1 << 20  # never dereferenced by the planner
    for i, st in enumerate((w.num_cams * Hf * Wf * Cc, Hf * Wf * Cc, 1, Wf * Cc, Cc)):  # channels-last
        a.feats_stride[i] = st
    a.masks, a.masks_dtype = 1 << 20, abi.MASK_PACKED
    words = (w.mask_hw[0] + 2) * lib.mvsgi_mask_pitch_words(w.mask_hw[1]) * 2
    a.masks_stride[0], a.masks_stride[1] = 0, words
    a.rig = C.pointer(rig)
    need = lib.mvsgi_sweep_tap_table_bytes(C.byref(a))
    pairs = w.cv_hw[0] * w.cv_hw[1] * len(w.dist_list)
    assert need % 256 == 0 and pairs * 44 <= need <= int(pairs * 44 * 1.15), need
    # per-frame masks: the validity bits of the records change with the frame
    a.masks_stride[0] = w.num_cams * words
    assert lib.mvsgi_sweep_tap_table_bytes(C.byref(a)) == 0
    a.vol, a.tap_table, a.tap_table_bytes, a.tap_table_mode = 1 << 20, 1 << 20, need, abi.TAP_TABLE_USE
    assert lib.mvsgi_sweep_launch_count(C.byref(a)) == abi.E_UNSUPPORTED
    a.masks_stride[0] = 0
    assert lib.mvsgi_sweep_launch_count(C.byref(a)) >= 1
    a.tap_table_bytes = need - 1
    assert lib.mvsgi_sweep_launch_count(C.byref(a)) < 0


# ------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvs_gi_b200 import abi

    abi.load_library()
    return torch.device("cuda:0")


def _nhwc(f):
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


def _inputs(dev, name, B, kw, seed=5, nhwc=True):
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.WORKLOADS[name]
    if kw:
        w = synthetic.small(w, **kw)
    std = w.aggregate == "std"
    g = torch.Generator(device=dev).manual_seed(seed)
    feats = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=g)
    if nhwc:
        feats = _nhwc(feats)
    blob = ops.rig_to_blob(w.rig().to_struct())
    bits = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]) if std else None
    agg = abi.AGG_STD if std else abi.AGG_CONCAT

    def mk(f=feats, channels_last=True, masks=bits, mask_hw=tuple(w.mask_hw) if std else None):
        return ops.SweepCall(f, None, None, masks, blob, agg, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                             channels_last=channels_last, mask_hw=mask_hw)

    return w, feats, mk


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,kw,balanced,nhwc,cl", [
    ("G16V", 1, {}, False, True, True),       # one frame per pass
    ("G16V", 5, {}, True, True, True),        # BUILD 2 frames per pass (ragged last pass), USE 1 frame per pass; balanced runs
    ("G16V", 17, {}, True, True, True),       # 4 frames per pass
    ("G16V", 7, {}, False, True, True),       # kept table: 4 frames per pass from 4 frames on, ragged last pass
    ("G16V", 6, {}, False, True, True),       # ... 2 frames per pass where that re-sweeps fewer frames
    ("4CAM", 4, {}, False, True, True),
    ("G16V", 2, {}, False, False, False),     # the reference layouts: NCHW features (staged), NCDHW volume
    ("4CAM", 6, {}, True, True, True),        # two waves of runs per CTA: the mbarrier phase flips between items
    ("4CAM", 6, {}, False, True, True),       # ... without the table of runs: the later items come from the queue
    ("HIGHD", 2, {}, False, True, True),      # 64 candidates in chunks of 16: 16 items per CTA, drawn from the queue
    ("G8", 3, {}, True, True, True),          # concat builder
    ("G16V_DS", 2, {}, False, True, True),    # DoubleSphere projection
    ("G16VV_C48", 2, {}, False, True, True),  # three channel groups
    ("G16V", 2, dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 77), n_cands=16), False, True, True),  # ragged
    ("G16V", 3, dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(3, 8), n_cands=5), False, True, True),  # fewer items than CTAs
])
def test_kept_tap_table_gives_the_same_volume_bit_for_bit(dev, name, B, kw, balanced, nhwc, cl):
    from mvs_gi_b200 import abi, ops

    w, feats, mk0 = _inputs(dev, name, B, kw, nhwc=nhwc)
    mk = lambda f=feats: mk0(f, channels_last=cl)
    plain, kept = mk(), mk()
    order = None
    if balanced:
        order = ops.balanced_octet_order(kept)
        assert order is not None
        kept.set_octet_order(order)
    need = kept.tap_table_bytes()
    assert need > 0 and need % 256 == 0
    plain.run()
    table = ops.build_tap_table(kept)  # one BUILD launch: builds as usual, stores the records
    torch.cuda.synchronize()
    assert table is not None and table.numel() == need and kept.args.tap_table_mode == abi.TAP_TABLE_USE
    assert torch.equal(plain.vol.view(torch.int32), kept.vol.view(torch.int32)), "BUILD launch"
    assert int(table.view(torch.int32).ne(0).sum()) > 0
    kept.vol.fill_(float("nan"))
    kept.run()  # USE
    torch.cuda.synchronize()
    assert torch.equal(plain.vol.view(torch.int32), kept.vol.view(torch.int32)), "USE launch"
    # the table belongs to the problem, not to the tensors: another call with other features of the same layout
    g = torch.Generator(device=dev).manual_seed(99)
    f2 = torch.randn(feats.shape, device=dev, generator=g)
    f2 = _nhwc(f2) if nhwc else f2
    p2, k2 = mk(f2), mk(f2)
    k2.set_octet_order(order)
    k2.set_tap_table(table)
    p2.run()
    k2.run()
    torch.cuda.synchronize()
    assert torch.equal(p2.vol.view(torch.int32), k2.vol.view(torch.int32)), "USE launch, other features"
    assert not torch.equal(p2.vol.view(torch.int32), plain.vol.view(torch.int32))


@pytest.mark.gpu
def test_tap_table_is_refused_where_the_records_change_per_frame(dev):
    from mvs_gi_b200 import abi, ops, synthetic

    w, feats, mk = _inputs(dev, "G16V", 3, dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=8))
    # per-frame masks (the training interface): nothing to keep
    m = torch.from_numpy(synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), 3, 1)).to(dev)
    per_frame = mk(masks=m, mask_hw=None)
    assert per_frame.tap_table_bytes() == 0
    assert ops.build_tap_table(per_frame) is None
    per_frame.set_tap_table(torch.zeros(1 << 20, dtype=torch.uint8, device=dev), abi.TAP_TABLE_BUILD)
    with pytest.raises(RuntimeError):
        per_frame.run()
    # a buffer smaller than mvsgi_sweep_tap_table_bytes, a misaligned one and an unknown mode
    call = mk()
    need = call.tap_table_bytes()
    assert need > 0
    call.set_tap_table(torch.zeros(need - 256, dtype=torch.uint8, device=dev), abi.TAP_TABLE_BUILD)
    with pytest.raises(RuntimeError):
        call.run()
    call.set_tap_table(torch.zeros(need + 256, dtype=torch.uint8, device=dev)[16:], abi.TAP_TABLE_BUILD)
    with pytest.raises(RuntimeError):
        call.run()
    call.set_tap_table(torch.zeros(need, dtype=torch.uint8, device=dev), 7)
    with pytest.raises(RuntimeError):
        call.run()
    call.set_tap_table(None)
    call.run()
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_module_builds_the_tap_table_once_and_reuses_it(dev):
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16)
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]
    g = torch.Generator(device=dev).manual_seed(3)
    feats = [_nhwc(torch.randn((2, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=g)) for _ in range(2)]

    def module(keep):
        m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval()
        m.attach_rig(w.rig()).set_static_masks(masks).set_vol_channels_last(True)
        m.keep_tap_table = keep
        return m

    kept, plain = module(True), module(False)
    with torch.no_grad():
        v0 = kept.sweep(feats[0], None, None, None)
        table = kept._order_cache[2]
        assert table is not None and table.dtype == torch.uint8
        v1 = kept.sweep(feats[1], None, None, None)  # USE
        assert kept._order_cache[2] is table
        r0, r1 = plain.sweep(feats[0], None, None, None), plain.sweep(feats[1], None, None, None)
        assert plain._order_cache[2] is None
    torch.cuda.synchronize()
    assert torch.equal(v0.view(torch.int32), r0.view(torch.int32))
    assert torch.equal(v1.view(torch.int32), r1.view(torch.int32))
    # new static masks: the records hold validity bits, so the table is rebuilt
    masks2 = masks.clone()
    masks2[..., : masks.shape[-2] // 2, :] = 0
    kept.set_static_masks(masks2)
    plain.set_static_masks(masks2)
    with torch.no_grad():
        v2, r2 = kept.sweep(feats[0], None, None, None), plain.sweep(feats[0], None, None, None)
    torch.cuda.synchronize()
    assert kept._order_cache[2] is not table
    assert torch.equal(v2.view(torch.int32), r2.view(torch.int32))
    assert not torch.equal(v2.view(torch.int32), v0.view(torch.int32))
    # training path: gradients flow through the op with a table attached
    f = feats[0].clone().requires_grad_(True)
    kept.sweep(f, None, None, None).sum().backward()
    assert f.grad is not None and torch.isfinite(f.grad).all()


@pytest.mark.gpu
def test_item_queue_of_a_kept_table_replays_from_a_cuda_graph(dev):
    """More items than CTAs + a kept table: the CTAs draw their later items from a counter in the call's workspace that
    mvsgi_sweep_f32 zeroes on the stream (a memset node in a captured graph) — every replay sweeps every item once."""
    from mvs_gi_b200 import ops

    w, feats, mk = _inputs(dev, "HIGHD", 1, {})
    plain, kept = mk(), mk()
    plain.run()
    ops.build_tap_table(kept)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        kept.run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        kept.run()
    for _ in range(3):
        kept.vol.fill_(float("nan"))
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(plain.vol.view(torch.int32), kept.vol.view(torch.int32))
