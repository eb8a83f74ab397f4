"""Cost-balanced assignment of the output image's octets to the sweep's CTAs (MvsgiSweepArgs.octet_order).

With a static rig and static masks the work of every 8-pixel octet — the cameras its voxels will sample
(spherical_sweep_avg.py:92-111) — is known before the first frame.  `mvsgi_sweep_octet_cost` counts it on the device,
`ops.lpt_assign` deals the octets to the resident CTAs longest-first, and the sweep kernel follows the table instead
of its round-robin default.  The volume must not change by a bit: only which CTA computes which octet does."""
import numpy as np
import pytest
import torch


def test_lpt_assign_is_a_partition_within_capacity():
    from mvs_gi_b200 import ops

    rng = np.random.default_rng(3)
    for n, runs, units in [(3200, 296, 11), (17, 4, 5), (8, 8, 1), (5, 9, 1), (1, 1, 1), (100, 7, 15)]:
        cost = rng.integers(1, 2000, n)
        order = ops.lpt_assign(cost, runs, units)
        assert order.shape == (runs, units) and order.dtype == np.int32
        used = order[order >= 0]
        assert np.array_equal(np.sort(used), np.arange(n)), "every octet exactly once"
        # empty slots only at the end of a run
        for r in range(runs):
            k = int((order[r] >= 0).sum())
            assert np.all(order[r, :k] >= 0) and np.all(order[r, k:] == -1)
    # balance: a wide cost distribution (masked-out octets cost 1/40 of fully visible ones) ends within 5 % of the mean,
    # where the round-robin deal of the same costs spreads by more than 10 %
    cost = np.where(rng.random(3200) < 0.35, 24, rng.integers(400, 1500, 3200))
    order = ops.lpt_assign(cost, 296, 11)
    load = np.array([cost[o[o >= 0]].sum() for o in order])
    rr = np.array([cost[r::291].sum() for r in range(291)])
    assert load.max() / load.mean() < 1.05 and rr.max() / rr.mean() > 1.10
    with pytest.raises(ValueError):
        ops.lpt_assign(np.ones(10), 3, 3)


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


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,kw", [
    ("G16V", 4, {}),                                                     # full size, 4 frames per pass is off (B < 16)
    ("G16V", 17, {}),                                                    # 4 frames per pass, ragged last pass
    ("4CAM", 6, {}),
    ("G8", 3, {}),                                                       # concat builder: no masks
    ("G16V", 2, dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 77), n_cands=16)),   # ragged last octet
    ("G16V_DS", 2, dict(feat_hw=(32, 128), mask_hw=(152, 152), cv_hw=(12, 40), n_cands=7)),  # fewer octets than CTAs
])
def test_balanced_runs_give_the_same_volume_bit_for_bit(dev, name, B, kw):
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.WORKLOADS[name]
    if kw:
        w = synthetic.small(w, **kw)
    std = w.aggregate == "std"
    agg = abi.AGG_STD if std else abi.AGG_CONCAT
    g = torch.Generator(device=dev).manual_seed(11)
    feats = _nhwc(torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=g))
    blob = ops.rig_to_blob(w.rig().to_struct())
    bits = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]) if std else None
    mk = lambda: ops.SweepCall(feats, None, None, bits, blob, agg, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                               channels_last=True, mask_hw=tuple(w.mask_hw) if std else None)
    plain, bal = mk(), mk()
    order = ops.balanced_octet_order(bal)
    assert order is not None and order.dtype == torch.int32
    if name == "4CAM" and not kw:
        # four cameras at full size: the tap table of 1/296 of the image does not fit one pass; the table holds two
        # waves of runs, CTA c sweeps runs c and c + 296
        assert order.shape[0] == 2 * bal.lib.mvsgi_sweep_grid(__import__("ctypes").byref(bal.args))
    P = w.cv_hw[0] * w.cv_hw[1]
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o[o >= 0]), np.arange((P + 7) // 8))
    bal.set_octet_order(order)
    plain.run()
    bal.run()
    torch.cuda.synchronize()
    assert torch.equal(plain.vol.view(torch.int32), bal.vol.view(torch.int32))
    # a deliberately bad but valid table (everything in index order) is still the same volume
    runs, units = o.shape
    ident = np.full((runs, units), -1, np.int32)
    n_oct = (P + 7) // 8
    ident.reshape(-1)[:n_oct] = np.arange(n_oct)
    bal.set_octet_order(torch.from_numpy(ident).to(dev))
    bal.vol.zero_()
    bal.run()
    torch.cuda.synchronize()
    assert torch.equal(plain.vol.view(torch.int32), bal.vol.view(torch.int32))
    # a table with fewer slots than the image has octets is refused, not mis-swept
    if n_oct > 1:
        bal.set_octet_order(torch.zeros((1, 1), dtype=torch.int32, device=dev))
        with pytest.raises(RuntimeError):
            bal.run()


@pytest.mark.gpu
def test_octet_cost_counts_the_cameras_the_sweep_loads(dev):
    """cost = sum over candidates of 4 + 10 * (cameras the octet's warp loads).  Held against the validity volume the
    sweep itself reports (valid_out): a candidate with no valid voxel in the octet loads nothing, one with a valid voxel
    loads between 2 and CAM cameras."""
    import ctypes as C

    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(24, 80), n_cands=16)
    feats = torch.randn((1, w.num_cams, w.feat_chs, *w.feat_hw), device=dev)
    blob = ops.rig_to_blob(w.rig().to_struct())
    bits = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
    call = ops.SweepCall(feats, None, None, bits, blob, abi.AGG_STD, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                         want_valid=True, mask_hw=tuple(w.mask_hw))
    call.run()
    n_oct = w.cv_hw[0] * w.cv_hw[1] // 8
    cost = torch.empty((n_oct,), dtype=torch.int32, device=dev)
    abi.check(call.lib.mvsgi_sweep_octet_cost(C.byref(call.args), cost.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "mvsgi_sweep_octet_cost")
    torch.cuda.synchronize()
    valid = call.valid[0].to(torch.int64)                               # [D, Ho, Wo], 1 where more than one camera is valid
    D = valid.shape[0]
    live = (valid.reshape(D, n_oct, 8).sum(2) > 0).sum(0)              # candidates with a valid voxel, per octet
    c = cost.to(torch.int64)
    assert torch.all(c >= 4 * D + 20 * live) and torch.all(c <= 4 * D + 10 * w.num_cams * live)
    assert 0 < int((live == 0).sum()) < n_oct, "the case must hold both masked-out and visible octets"


@pytest.mark.gpu
def test_module_path_balances_once_and_matches(dev):
    """modules.SphericalSweepStdMasked with an attached rig and static masks computes the table at the first forward and
    reuses it; balance_runs = False gives the same volume."""
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16)
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval()
    m.attach_rig(w.rig())
    m.set_vol_channels_last(True)
    m.set_static_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
    m.balance_runs = True  # opt-in
    feats = _nhwc(torch.randn((3, w.num_cams, w.feat_chs, *w.feat_hw), device=dev))
    with torch.no_grad():
        v1 = m.sweep(feats, None, None, None)
        assert m._order_cache is not None and m._order_cache[1] is not None
        table = m._order_cache[1]
        v2 = m.sweep(feats, None, None, None)
        assert m._order_cache[1] is table
        m.balance_runs = False
        v3 = m.sweep(feats, None, None, None)
    assert torch.equal(v1, v2) and torch.equal(v1, v3)


def test_lpt_assign_gives_slow_runs_less():
    """slowdown: a run on an SM that needs 8 % longer per unit of cost gets about 8 % less cost."""
    from mvs_gi_b200 import ops

    rng = np.random.default_rng(5)
    cost = rng.integers(64, 560, 3200)
    slow = np.ones(296)
    slow[[7, 111, 290]] = 1.08
    order = ops.lpt_assign(cost, 296, 11, slow)
    used = order[order >= 0]
    assert np.array_equal(np.sort(used), np.arange(3200))
    load = np.array([cost[o[o >= 0]].sum() for o in order], float)
    t = load * slow
    assert t.max() / np.median(t) < 1.02
    assert np.all(load[[7, 111, 290]] < 0.95 * load.mean())
    with pytest.raises(ValueError):
        ops.lpt_assign(cost, 296, 11, np.ones(5))


@pytest.mark.gpu
def test_cta_times_and_calibrated_order(dev):
    """cta_times_out: one row per CTA (SM id, start <= end of build <= end); the calibrated table is a valid table (the
    same volume bit for bit) and the report says what was measured."""
    from mvs_gi_b200 import abi, ops, synthetic

    w = synthetic.WORKLOADS["G16V"]
    B = 8
    feats = _nhwc(torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev))
    blob = ops.rig_to_blob(w.rig().to_struct())
    bits = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
    mk = lambda: ops.SweepCall(feats, None, None, bits, blob, abi.AGG_STD, w.cv_hw[0], w.cv_hw[1], len(w.dist_list),
                               channels_last=True, mask_hw=tuple(w.mask_hw))
    plain, cal = mk(), mk()
    plain.run()
    sm, us = ops.cta_times(plain, launches=2)
    assert us.shape[0] == 2 and us.shape[1] == sm.shape[0] > 100
    assert np.all(us > 1.0) and np.all(us < 5e3) and sm.max() < 512 and len(np.unique(sm)) > 100
    assert plain.args.cta_times_out is None, "the timing buffer is detached again"
    order, report = ops.calibrated_octet_order(cal)
    assert order is not None and "kernel_us_balanced" in report and report["kernel_us"] <= report["kernel_us_balanced"]
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o[o >= 0]), np.arange(w.cv_hw[0] * w.cv_hw[1] // 8))
    cal.set_octet_order(order)
    cal.vol.zero_()
    cal.run()
    plain.run()
    torch.cuda.synchronize()
    assert torch.equal(plain.vol.view(torch.int32), cal.vol.view(torch.int32))
