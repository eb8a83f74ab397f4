"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libmvsgi_b200.so (via mvs_gi_b200.ops / modules) and is compared with

  1. the golden vectors generated from the reference's own modules (tests/golden/, oracle/gen_golden.py),
  2. the CPU oracle (oracle/mvsgi_oracle.c) on seeded inputs the oracle finishes in seconds,
  3. size-independent properties at BASELINE.json's full sizes.

Pass criteria (SURVEY.md §8c / north_star): validity mask (= zero pattern of vol) and arg-max candidate bit-exact;
cost volume and inv_dist within 1e-5 * max|ref| in fp32.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: 1e-5 relative (normalised by max|ref|)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvs_gi_b200 import abi

    abi.load_library()  # fail loudly if the CUDA extension is missing
    return torch.device("cuda:0")


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def relmax(a, ref):
    return float(np.abs(a - ref).max() / max(float(np.abs(ref).max()), 1e-30))


# ------------------------------------------------------------------ 1. golden vectors from the reference modules
@pytest.mark.parametrize("name", ["std3", "std4"])
def test_golden_sweep_std(golden, dev, name):
    from mvs_gi_b200 import modules

    g = golden(f"sweep_{name}")
    B, CAM, C = g["feats"].shape[:3]
    m = modules.SphericalSweepStdMasked(CAM, C, 3).to(dev).eval()
    with torch.no_grad():
        vol, valid = m.sweep_with_valid(t(g["feats"], dev), t(g["grids"], dev), t(g["grid_masks"], dev),
                                        t(g["masks"], dev))
    vol, valid, ref = vol.cpu().numpy(), valid.cpu().numpy(), g["vol"]
    assert np.array_equal(vol == 0, ref == 0), "zero pattern / validity mask must be bit-exact"
    assert np.array_equal(np.broadcast_to(valid[:, None], ref.shape), ref != 0)
    assert relmax(vol, ref) <= TOL


def test_golden_sweep_concat(golden, dev):
    from mvs_gi_b200 import modules

    g = golden("sweep_concat3")
    B, CAM, C = g["feats"].shape[:3]
    m = modules.SphericalSweep(CAM, CAM * C, 3).to(dev).eval()
    with torch.no_grad():
        vol = m.sweep(t(g["feats"], dev), t(g["grids"], dev), None)
    assert vol.shape == g["vol"].shape
    assert relmax(vol.cpu().numpy(), g["vol"]) <= TOL


@pytest.mark.parametrize("name", ["sa_d16_s2", "sa_d8_s0", "sa_d8_s2_nopre", "sa_d5_s3"])
def test_golden_softargmin(golden, dev, name):
    from mvs_gi_b200 import modules

    g = golden(f"softargmin_{name}")
    reg = modules.DistanceRegressorWithFixedCandidates(96, list(g["dist"]), float(g["scale"]),
                                                       bool(g["pre_interp"])).to(dev)
    assert np.array_equal(reg.inv_dist_idx.cpu().numpy().reshape(-1), g["inv_dist_idx"])
    with torch.no_grad():
        inv, nc, am = reg.regress(t(g["costs"], dev), want_norm_costs=True, want_argmax=True)
        inv2, nc2 = reg(t(g["costs"], dev))
    assert torch.equal(inv, inv2) and torch.equal(nc, nc2)
    assert inv.shape == g["inv_dist"].shape and nc.shape == g["norm_costs"].shape
    assert np.array_equal(am.cpu().numpy(), g["argmax"]), "arg-max candidate must be bit-exact"
    assert np.array_equal(nc.argmax(1).cpu().numpy(), g["argmax"])
    assert relmax(inv.cpu().numpy(), g["inv_dist"]) <= TOL
    assert np.abs(nc.cpu().numpy() - g["norm_costs"]).max() <= TOL


# ------------------------------------------------------------------ 2. against the CPU oracle
def _case(name, B, seed, **small_kw):
    from mvs_gi_b200 import synthetic

    w = synthetic.small(synthetic.WORKLOADS[name], **small_kw)
    feats = synthetic.make_feats(w, B, seed)
    base = synthetic.fisheye_footprint_masks(w)
    masks = synthetic.erase_rectangles(base, B, seed + 1)
    return w, feats, base, masks


@pytest.mark.parametrize("name,kw", [
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16)),
    ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16)),
    ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6)),
    ("G16V_DS", dict(feat_hw=(32, 128), cv_hw=(20, 80), n_cands=5)),
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3)),  # ragged: nothing divides a tile
])
def test_fused_projection_vs_oracle(dev, name, kw):
    """Fused in-kernel projection.  The oracle is fed the kernel's own grids (mvsgi_project_only), so masks are
    bit-exact by construction of the comparison, and the grids themselves are bounded separately below."""
    from mvs_gi_b200 import modules, ops
    from oracle import oracle as orc

    B = 3
    w, feats, _, masks = _case(name, B, 11, **kw)
    rig = w.rig()
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    with torch.no_grad():
        vol, valid = m.sweep_with_valid(t(feats, dev), None, None, t(masks, dev))
        grids, gm = ops.project_only(rig.to_struct(), dev)
    ref, ref_valid = orc.sweep_std(feats, grids.cpu().numpy()[None], gm.cpu().numpy()[None], masks, return_valid=True)
    vol = vol.cpu().numpy()
    assert np.array_equal(valid.cpu().numpy(), ref_valid)
    assert np.array_equal(vol == 0, ref == 0)
    assert 0.02 < ref_valid.mean() < 0.999
    assert relmax(vol, ref) <= TOL


@pytest.mark.parametrize("name", ["G16V", "4CAM", "G16V_DS"])
def test_projection_accuracy(dev, name):
    """Kernel grids vs the fp32 oracle (reference op order) and vs float64: error no larger than the reference's own
    distance from float64 (SURVEY.md H2, 1.1e-5 normalised), validity bits equal except ulp-boundary samples."""
    from mvs_gi_b200 import ops, synthetic
    from oracle import oracle as orc

    w = synthetic.small(synthetic.WORKLOADS[name], cv_hw=(40, 160))
    if w.camera == "erp":
        w.mask_hw = synthetic.WORKLOADS[name].mask_hw
    rs = w.rig().to_struct()
    grids, gm = ops.project_only(rs, dev)
    grids, gm = grids.cpu().numpy(), gm.cpu().numpy()[..., 0]
    g32, m32 = orc.make_grids(rs)
    g64, m64 = orc.make_grids_f64(rs)
    both = gm & m64
    err64 = np.abs(grids - g64)[both].max()
    err32 = np.abs(grids - g32)[gm & m32].max()
    flips = int((gm != m32).sum())
    print(f"{name}: |kernel-f64|={err64:.2e} |kernel-oracle32|={err32:.2e} flips={flips}/{gm.size}")
    assert err64 <= 1.1e-5
    assert flips <= max(2, gm.size // 5000)


def test_supplied_grids_vs_oracle_per_frame_everything(dev):
    """Reference interface: per-frame grids, grid_masks and masks (nothing shared)."""
    from mvs_gi_b200 import modules
    from oracle import oracle as orc

    B = 3
    w, feats, _, masks = _case("G16V", B, 23, feat_hw=(24, 96), mask_hw=(96, 384), cv_hw=(12, 72), n_cands=7)
    g0, m0 = orc.make_grids(w.rig().to_struct())
    gen = np.random.default_rng(5)
    grids = np.repeat(g0[None], B, 0) + gen.normal(0, 0.02, (B, *g0.shape)).astype(np.float32)
    gmask = np.repeat(m0[None], B, 0) & (gen.random((B, *m0.shape)) > 0.1)
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval()
    with torch.no_grad():
        vol = m.sweep(t(feats, dev), t(grids, dev), t(gmask[..., None], dev), t(masks, dev)).cpu().numpy()
    ref = orc.sweep_std(feats, grids, gmask, masks)
    assert np.array_equal(vol == 0, ref == 0)
    assert relmax(vol, ref) <= TOL


def test_layouts_and_sharing_are_equivalent(dev):
    """NCHW vs channels-last features, bool vs float masks, shared vs replicated masks/grids: identical bits out."""
    from mvs_gi_b200 import modules
    from oracle import oracle as orc

    B = 2
    w, feats, base, _ = _case("G16V", B, 31, feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=4)
    g0, m0 = orc.make_grids(w.rig().to_struct())
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval()
    f = t(feats, dev)
    grids1, gm1, masks1 = t(g0[None], dev), t(m0[None, ..., None], dev), t(base[None], dev)
    with torch.no_grad():
        a = m.sweep(f, grids1.expand(B, *grids1.shape[1:]), gm1.expand(B, *gm1.shape[1:]),
                    masks1.expand(B, *masks1.shape[1:]))
        b = m.sweep(f, grids1.repeat(B, 1, 1, 1, 1, 1), gm1.repeat(B, 1, 1, 1, 1, 1), masks1.repeat(B, 1, 1, 1, 1))
        f_cl = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)  # channels-last memory, same logical shape
        assert f_cl.stride(2) == 1
        c = m.sweep(f_cl, grids1, gm1, masks1)
        d = m.sweep(f, grids1, gm1, masks1.bool())
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, d)
    ref = orc.sweep_std(feats, g0[None], m0[None], base[None])
    assert np.array_equal(a.cpu().numpy() == 0, ref == 0)
    assert relmax(a.cpu().numpy(), ref) <= TOL


def test_concat_vs_oracle_and_edges(dev):
    from mvs_gi_b200 import modules, synthetic
    from oracle import oracle as orc

    B = 2
    w = synthetic.small(synthetic.WORKLOADS["G8"], feat_hw=(16, 64), cv_hw=(10, 40))
    feats = synthetic.make_feats(w, B, 41)
    gen = np.random.default_rng(6)
    grids = gen.uniform(-1.3, 1.3, (B, w.num_cams, w.num_cands, 10, 40, 2)).astype(np.float32)
    grids[0, 0, 0, 0, :6] = [[-1, -1], [1, 1], [-1 - 1e-6, 0], [0, 1 + 1e-6], [-7, 0.3], [0.1, 9]]
    grids[0, 1, 0, 0, 0] = [np.nan, 0.0]  # NaN coordinate: zero taps here (documented deviation: reference yields NaN)
    m = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval()
    with torch.no_grad():
        vol = m.sweep(t(feats, dev), t(grids, dev), None).cpu().numpy()
    grids_o = grids.copy()
    grids_o[0, 1, 0, 0, 0] = [5.0, 0.0]  # oracle: a far-outside point is also exactly zero
    ref = orc.sweep_concat(feats, grids_o)
    assert relmax(vol, ref) <= TOL
    assert np.all(vol[0, w.feat_chs:2 * w.feat_chs, 0, 0, 0] == 0)


@pytest.mark.parametrize("D,scale", [(16, 2), (32, 2), (64, 2), (7, 0), (24, 3)])
def test_softargmin_vs_oracle(dev, D, scale):
    """D > 16 uses the warp-shuffle split-D variant."""
    from mvs_gi_b200 import calibration as cal
    from mvs_gi_b200 import modules
    from oracle import oracle as orc

    dist = cal.dist_inverse_spaced(D)
    gen = np.random.default_rng(D)
    costs = (gen.standard_normal((2, 1, D, 21, 45)) * 4).astype(np.float32)
    reg = modules.DistanceRegressorWithFixedCandidates(96, dist, scale, scale > 0).to(dev)
    with torch.no_grad():
        inv, nc, am = reg.regress(t(costs, dev), want_norm_costs=True, want_argmax=True)
    r_inv, r_nc, r_am = orc.softargmin(costs, reg.inv_dist_idx.cpu().numpy(), scale, scale > 0)
    assert np.array_equal(am.cpu().numpy(), r_am)
    assert relmax(inv.cpu().numpy(), r_inv) <= TOL
    assert np.abs(nc.cpu().numpy() - r_nc).max() <= TOL


def test_candidate_slices_and_peer_stores_on_one_gpu(dev):
    """Candidate-sharded entry of the C ABI on a single GPU: two calls that each sweep a slice of the candidate list
    into one full volume reproduce the un-sharded sweep bit for bit, and a "peer" volume (here a second local buffer
    standing in for another GPU's mapped memory) receives the identical stores."""
    from mvs_gi_b200 import abi, modules, ops, sharding

    B = 2
    w, feats, _, masks = _case("G16V", B, 51, feat_hw=(24, 96), mask_hw=(96, 384), cv_hw=(12, 72), n_cands=7)
    rig = w.rig()
    f, mk = t(feats, dev), t(masks, dev)
    full = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    with torch.no_grad():
        ref = full.sweep(f, None, None, mk)
    vol = torch.full_like(ref, float("nan"))
    peer = torch.full_like(ref, float("nan"))
    for r in range(2):
        d0, d1 = sharding.balanced_range(w.num_cands, r, 2)
        blob = ops.rig_to_blob(sharding.sub_rig(rig, d0, d1).to_struct())
        call = ops.SweepCall(f, None, None, mk, blob, abi.AGG_STD, *w.cv_hw, d1 - d0, out=vol, cand_offset=d0,
                             peer_ptrs=[peer.data_ptr()])
        call.run()
    assert torch.equal(vol, ref) and torch.equal(peer, ref)
    # concat builder, supplied grids sliced along D
    from oracle import oracle as orc
    g0, _ = orc.make_grids(rig.to_struct())
    grids = t(g0[None], dev)
    cat = modules.SphericalSweep(w.num_cams, w.feat_chs * w.num_cams, 3).to(dev).eval()
    with torch.no_grad():
        ref_c = cat.sweep(f, grids, None)
    vol_c = torch.full_like(ref_c, float("nan"))
    for r in range(3):
        d0, d1 = sharding.balanced_range(w.num_cands, r, 3)
        ops.SweepCall(f, grids[:, :, d0:d1], None, None, None, abi.AGG_CONCAT, *w.cv_hw, d1 - d0, out=vol_c,
                      cand_offset=d0).run()
    assert torch.equal(vol_c, ref_c)
    with pytest.raises(RuntimeError, match="fit"):
        ops.SweepCall(f, None, None, mk, ops.rig_to_blob(rig.to_struct()), abi.AGG_STD, *w.cv_hw, w.num_cands, out=vol,
                      cand_offset=1)


def test_channels_last_volume_is_the_same_volume(dev):
    """MVSGI_FLAG_VOL_CHANNELS_LAST: the [B,C,D,Ho,Wo] volume written with channels_last_3d
    strides holds bit-identical values (std 3/4 cameras incl. C=12 and C=48, concat, candidate slices + peer stores),
    and post_vol (cuDNN Conv3d + BN + LeakyReLU, spherical_sweep_avg.py:165) gives the reference-layout result."""
    from mvs_gi_b200 import abi, modules, ops, sharding, synthetic
    from oracle import oracle as orc

    for name, kw, chs in [("G16V", dict(feat_hw=(24, 96), mask_hw=(96, 384), cv_hw=(12, 72), n_cands=7), None),
                          ("4CAM", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3), 12),
                          ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), None)]:
        w, feats, _, masks = _case(name, 2, 61, **kw)
        if chs is not None:
            feats = np.ascontiguousarray(feats[:, :, :chs])
        C = feats.shape[2]
        rig = w.rig()
        f, mk = t(feats, dev), t(masks, dev)
        m = modules.SphericalSweepStdMasked(w.num_cams, C, 3).to(dev).eval().attach_rig(rig)
        with torch.no_grad():
            ref = m.sweep(f, None, None, mk)
            ref_post = m(f, None, None, mk)
            m.set_vol_channels_last(True)
            cl = m.sweep(f, None, None, mk)
            cl_post = m(f, None, None, mk)
        assert cl.shape == ref.shape and cl.is_contiguous(memory_format=torch.channels_last_3d)
        assert cl.stride(1) == 1 and not cl.is_contiguous()
        assert torch.equal(cl, ref)
        assert relmax(cl_post.cpu().numpy(), ref_post.cpu().numpy()) <= 1e-4  # cuDNN picks another algorithm
        # candidate slices + a stand-in peer volume, channels-last
        vol = torch.full_like(cl, float("nan"))
        peer = torch.full_like(cl, float("nan"))
        assert vol.is_contiguous(memory_format=torch.channels_last_3d)
        for r in range(2):
            d0, d1 = sharding.balanced_range(w.num_cands, r, 2)
            blob = ops.rig_to_blob(sharding.sub_rig(rig, d0, d1).to_struct())
            ops.SweepCall(f, None, None, mk, blob, abi.AGG_STD, *w.cv_hw, d1 - d0, out=vol, cand_offset=d0,
                          peer_ptrs=[peer.data_ptr()], channels_last=True).run()
        assert torch.equal(vol, ref) and torch.equal(peer, ref)
        with pytest.raises(RuntimeError, match="channels_last_3d"):
            ops.SweepCall(f, None, None, mk, ops.rig_to_blob(rig.to_struct()), abi.AGG_STD, *w.cv_hw, w.num_cands,
                          out=torch.empty_like(ref), channels_last=True)
    # concat
    w = synthetic.small(synthetic.WORKLOADS["G8"], feat_hw=(16, 64), cv_hw=(10, 40))
    feats = synthetic.make_feats(w, 2, 62)
    g0, _ = orc.make_grids(w.rig().to_struct())
    cat = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval()
    with torch.no_grad():
        ref = cat.sweep(t(feats, dev), t(g0[None], dev), None)
        cl = cat.set_vol_channels_last(True).sweep(t(feats, dev), t(g0[None], dev), None)
    assert cl.stride(1) == 1 and torch.equal(cl, ref)
    # channels not a multiple of 4 cannot be stored as 16-byte vectors: loud error, no fallback
    w, feats, _, masks = _case("G16V", 1, 63, feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=2)
    m = modules.SphericalSweepStdMasked(w.num_cams, 6, 3).to(dev).eval().attach_rig(w.rig()).set_vol_channels_last()
    with pytest.raises(RuntimeError, match="shape"):
        m.sweep(t(np.ascontiguousarray(feats[:, :, :6]), dev), None, None, t(masks, dev))


def test_walking_quads_experiment_is_bit_identical(dev, monkeypatch):
    """sweep_walk.cuh (MVSGI_WALK=1, not the default: it measures slower) must still be the same function:
    bit-identical to the default kernel for std / concat, 3 and 4 cameras, C=48, ragged sizes, per-frame masks."""
    from mvs_gi_b200 import modules, synthetic
    from oracle import oracle as orc

    for name, kw in [("G16V", dict(feat_hw=(24, 96), mask_hw=(96, 384), cv_hw=(12, 72), n_cands=7)),
                     ("4CAM", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3)),
                     ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=16))]:
        w, feats, _, masks = _case(name, 3, 71, **kw)
        m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(w.rig())
        m.set_vol_channels_last(True)
        with torch.no_grad():
            monkeypatch.setenv("MVSGI_WALK", "0")
            ref, ref_valid = m.sweep_with_valid(t(feats, dev), None, None, t(masks, dev))
            monkeypatch.setenv("MVSGI_WALK", "1")
            for segs in ("", "3", "32"):
                monkeypatch.setenv("MVSGI_WALK_SEGS", segs)
                got, got_valid = m.sweep_with_valid(t(feats, dev), None, None, t(masks, dev))
                assert torch.equal(got, ref) and torch.equal(got_valid, ref_valid), (name, segs)
    monkeypatch.setenv("MVSGI_WALK_SEGS", "")
    w = synthetic.small(synthetic.WORKLOADS["G8"], feat_hw=(16, 64), cv_hw=(10, 40))
    feats = synthetic.make_feats(w, 2, 72)
    g0, _ = orc.make_grids(w.rig().to_struct())
    grids = np.repeat(g0[None], 2, 0)
    grids[1] += 0.01  # per-frame grids: the table is rebuilt per frame
    cat = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval().set_vol_channels_last(True)
    with torch.no_grad():
        monkeypatch.setenv("MVSGI_WALK", "0")
        ref = cat.sweep(t(feats, dev), t(grids, dev), None)
        monkeypatch.setenv("MVSGI_WALK", "1")
        got = cat.sweep(t(feats, dev), t(grids, dev), None)
    assert torch.equal(got, ref)


def test_bf16_staged_features_stated_tolerance(dev):
    """MVSGI_FLAG_FEATS_BF16 (SURVEY.md §8c criterion 4): the kernel must equal the fp32 oracle run on
    bfloat16-rounded features to 1e-5 (same arithmetic, rounded inputs), keep the validity mask / zero pattern
    bit-exact, and stay within 1e-2 * max|ref| of the un-rounded fp32 volume.  Both volume layouts, std and concat."""
    from mvs_gi_b200 import modules, ops, synthetic
    from oracle import oracle as orc

    BF16_TOL = 1e-2
    for name, kw in [("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16)),
                     ("4CAM", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3)),
                     ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6))]:
        w, feats, _, masks = _case(name, 2, 81, **kw)
        rig = w.rig()
        m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
        with torch.no_grad():
            full = m.sweep(t(feats, dev), None, None, t(masks, dev))
            m.set_feats_bf16(True)
            got = m.sweep(t(feats, dev), None, None, t(masks, dev))
            got_cl = m.set_vol_channels_last(True).sweep(t(feats, dev), None, None, t(masks, dev))
            grids, gm = ops.project_only(rig.to_struct(), dev)
        assert torch.equal(got, got_cl)
        g, gmn = grids.cpu().numpy()[None], gm.cpu().numpy()[None]
        ref_rounded = orc.sweep_std(orc.bf16_round(feats), g, gmn, masks)
        ref = orc.sweep_std(feats, g, gmn, masks)
        v = got.cpu().numpy()
        assert np.array_equal(v == 0, ref == 0), "validity / zero pattern must not depend on the staging dtype"
        assert relmax(v, ref_rounded) <= TOL
        assert relmax(v, ref) <= BF16_TOL
        assert relmax(v, full.cpu().numpy()) <= BF16_TOL and not torch.equal(got, full)
        with pytest.raises(RuntimeError, match="unsupported"):
            m.sweep_with_valid(t(feats, dev), None, None, t(masks, dev))
    w = synthetic.small(synthetic.WORKLOADS["G8"], feat_hw=(16, 64), cv_hw=(10, 40))
    feats = synthetic.make_feats(w, 2, 82)
    g0, _ = orc.make_grids(w.rig().to_struct())
    cat = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval().set_feats_bf16(True)
    with torch.no_grad():
        got = cat.sweep(t(feats, dev), t(g0[None], dev), None).cpu().numpy()
    assert relmax(got, orc.sweep_concat(orc.bf16_round(feats), g0[None])) <= TOL
    assert relmax(got, orc.sweep_concat(feats, g0[None])) <= BF16_TOL


def test_cuda_graph_capture_and_replay(dev):
    """The C ABI neither allocates nor synchronises, so a whole step (staging, mask packing, sweep, soft-argmin) and a
    backward call capture into a CUDA graph; replays on new input values reproduce the eager results bit for bit."""
    from mvs_gi_b200 import abi, modules, ops

    w, feats, _, masks = _case("G16V", 2, 91, feat_hw=(24, 96), mask_hw=(96, 384), cv_hw=(12, 72), n_cands=7)
    rig = w.rig()
    blob = ops.rig_to_blob(rig.to_struct())
    f, mk = t(feats, dev), t(masks, dev)
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, 2, True).to(dev)
    costs = torch.randn((2, 1, w.num_cands, *w.cv_hw), device=dev)
    call = ops.SweepCall(f, None, None, mk, blob, abi.AGG_STD, *w.cv_hw, w.num_cands, channels_last=True)
    gvol = torch.randn_like(call.vol)

    def step():
        call.run()
        inv, _, _ = ops.softargmin_op(costs, reg.inv_dist_idx, 24, 144, 0.5, False, False)
        gf = ops.sweep_backward_op(gvol, f, None, None, mk, blob, abi.AGG_CONCAT if False else abi.AGG_STD, *w.cv_hw,
                                   w.num_cands)
        return inv, gf

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_inv, g_gf = step()
    for seed in (1, 2):
        torch.manual_seed(seed)
        f.copy_(torch.randn_like(f))
        costs.copy_(torch.randn_like(costs))
        graph.replay()
        torch.cuda.synchronize()
        got_vol, got_inv, got_gf = call.vol.clone(), g_inv.clone(), g_gf.clone()
        inv, gf = step()
        assert torch.equal(call.vol, got_vol) and torch.equal(inv, got_inv)
        # fp32 atomics: the gradient is order-dependent in the last bits
        assert float((gf - got_gf).abs().max()) <= 1e-5 * float(gf.abs().max())


def test_pack_masks_public_entry_layout(dev):
    """mvsgi_pack_masks as a caller of the C ABI sees it: the documented row-pair bit layout, the negative/NaN flag and
    the alignment check (include/mvsgi_b200.h)."""
    import ctypes as C

    from mvs_gi_b200 import abi

    lib = abi.load_library()
    Hm, Wm = 5, 70
    pitch = int(lib.mvsgi_mask_pitch_words(Wm))
    gen = np.random.default_rng(12)
    m = (gen.uniform(0, 1, (2, Hm, Wm)) > 0.5).astype(np.float32)
    masks = t(m, dev)
    bits = torch.full((2 * (Hm + 2) * pitch * 2 + 2,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def pack(src, out_ptr):
        return lib.mvsgi_pack_masks(src.data_ptr(), abi.MASK_F32, 2, Hm * Wm, 1, 0, Wm, 1, Hm, Wm, out_ptr,
                                    flags.data_ptr(), st)

    assert pack(masks, bits.data_ptr()) == 0
    torch.cuda.synchronize()
    w = bits.cpu().numpy().view(np.uint32)[:2 * (Hm + 2) * pitch * 2].reshape(2, Hm + 2, pitch, 2)
    padded = np.zeros((2, Hm + 3, pitch * 32), np.uint8)  # padded rows 0 .. Hm+2, bit 32 + x of a row is pixel x
    padded[:, 1:Hm + 1, 32:32 + Wm] = m.astype(np.uint8)
    for n in range(2):
        for Y in range(Hm + 2):
            for r in range(2):
                row_bits = np.unpackbits(w[n, Y, :, r].astype("<u4").view(np.uint8), bitorder="little")
                assert np.array_equal(row_bits, padded[n, Y + r]), (n, Y, r)
    assert int(flags.item()) == 0
    bad = masks.clone()
    bad[1, 2, 3] = -1.0
    assert pack(bad, bits.data_ptr()) == 0
    torch.cuda.synchronize()
    assert int(flags.item()) & 1
    assert pack(masks, bits.data_ptr() + 4) == -3  # MVSGI_E_STRIDE: 8-byte alignment


def test_candidate_sharded_sweep_two_gpus(dev):
    """NCCL all-gather variant and the fused peer-store variant on 2 GPUs of one box (skipped on a 1-GPU box)."""
    import json
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "dshard_bench.py"), "--workload",
           "G16V", "--batch", "2", "--steps", "2", "--small"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["bit_identical_to_unsharded_on_all_ranks"] is True


def test_errors_are_loud(dev):
    from mvs_gi_b200 import modules

    m = modules.SphericalSweepStdMasked(3, 16, 3).to(dev).eval()
    feats = torch.zeros(1, 3, 16, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.sweep(feats, torch.zeros(1, 3, 2, 4, 4, 2), torch.ones(1, 3, 2, 4, 4, 1, dtype=torch.bool),
                torch.ones(1, 3, 1, 8, 8))
    bad = modules.SphericalSweepStdMasked(5, 16, 3).to(dev).eval()  # more cameras than MVSGI_MAX_CAMS
    with pytest.raises(RuntimeError, match="shape"):
        bad.sweep(torch.zeros(1, 5, 16, 8, 8, device=dev), torch.zeros(1, 5, 2, 4, 4, 2, device=dev),
                  torch.ones(1, 5, 2, 4, 4, 1, dtype=torch.bool, device=dev), torch.ones(1, 5, 1, 8, 8, device=dev))
    # training path: gradients flow to feats (row f1); a wrong-shaped upstream gradient is an error, not a guess
    f = torch.zeros(1, 3, 16, 8, 8, device=dev, requires_grad=True)
    vol = m.sweep(f, torch.zeros(1, 3, 2, 4, 4, 2, device=dev),
                  torch.ones(1, 3, 2, 4, 4, 1, dtype=torch.bool, device=dev), torch.ones(1, 3, 1, 8, 8, device=dev))
    assert vol.requires_grad
    from mvs_gi_b200 import abi, ops
    with pytest.raises(RuntimeError, match="grad_vol"):
        ops.sweep_backward_op(torch.zeros(1, 16, 3, 4, 4, device=dev), f.detach(),
                              torch.zeros(1, 3, 2, 4, 4, 2, device=dev),
                              torch.ones(1, 3, 2, 4, 4, 1, dtype=torch.bool, device=dev),
                              torch.ones(1, 3, 1, 8, 8, device=dev), None, abi.AGG_STD, 4, 4, 2)


def test_camera_counts_candidate_limits_and_tiny_shapes(dev):
    """Edges of the supported range against the oracle: 1 and 2 cameras (std with one camera is all zero: n > 1 never
    holds, spherical_sweep_avg.py:106-111), the maximum of 64 in-kernel candidates, a single frame, a 1 x 1 cost
    volume, features of one pixel, and more candidates than MVSGI_MAX_CANDS (must be refused in fused mode)."""
    import copy

    from mvs_gi_b200 import calibration as cal
    from mvs_gi_b200 import modules, ops, synthetic
    from oracle import oracle as orc

    base = synthetic.small(synthetic.WORKLOADS["4CAM"], feat_hw=(12, 40), mask_hw=(24, 80), cv_hw=(6, 21), n_cands=5)
    for ncam in (1, 2):
        w = copy.copy(base)
        w.num_cams = ncam
        rig = w.rig()
        feats = synthetic.make_feats(w, 2, 100 + ncam)
        masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), 2, 7)
        g, gm = orc.make_grids(rig.to_struct())
        for cl in (False, True):
            m = modules.SphericalSweepStdMasked(ncam, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
            m.set_vol_channels_last(cl)
            with torch.no_grad():
                vol, valid = m.sweep_with_valid(t(feats, dev), None, None, t(masks, dev))
                kg, kgm = ops.project_only(rig.to_struct(), dev)
            ref, ref_valid = orc.sweep_std(feats, kg.cpu().numpy()[None], kgm.cpu().numpy()[None], masks,
                                           return_valid=True)
            assert np.array_equal(valid.cpu().numpy(), ref_valid)
            assert np.array_equal(vol.cpu().numpy() == 0, ref == 0)
            if ncam == 1:
                assert not ref_valid.any() and float(vol.abs().max()) == 0.0
            else:
                assert ref_valid.any() and relmax(vol.cpu().numpy(), ref) <= TOL
            cat = modules.SphericalSweep(ncam, ncam * w.feat_chs, 3).to(dev).eval().set_vol_channels_last(cl)
            with torch.no_grad():
                cv = cat.sweep(t(feats, dev), t(g[None], dev), None)
            assert relmax(cv.cpu().numpy(), orc.sweep_concat(feats, g[None])) <= TOL
    # 64 candidates evaluated in-kernel, one frame; then 65: refused
    w = copy.copy(base)
    w.num_cams = 3
    w.dist_list = cal.dist_inverse_spaced(64)
    rig = w.rig()
    feats = synthetic.make_feats(w, 1, 5)
    masks = synthetic.fisheye_footprint_masks(w)[None]
    m = modules.SphericalSweepStdMasked(3, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    with torch.no_grad():
        vol = m.sweep(t(feats, dev), None, None, t(masks, dev))
        kg, kgm = ops.project_only(rig.to_struct(), dev)
    ref = orc.sweep_std(feats, kg.cpu().numpy()[None], kgm.cpu().numpy()[None], masks)
    assert vol.shape[2] == 64 and np.array_equal(vol.cpu().numpy() == 0, ref == 0)
    assert relmax(vol.cpu().numpy(), ref) <= TOL
    with pytest.raises(ValueError, match="candidates"):
        cal.Rig(rig.cv_camera, rig.cameras, rig.poses_cv_cam, cal.dist_inverse_spaced(65))
    # supplied grids have no candidate limit: 70 candidates, 1 x 1 cost volume, 1 x 1 features
    gen = np.random.default_rng(9)
    feats1 = gen.standard_normal((2, 3, 16, 1, 1)).astype(np.float32)
    grids1 = gen.uniform(-1.2, 1.2, (2, 3, 70, 1, 1, 2)).astype(np.float32)
    cat = modules.SphericalSweep(3, 48, 3).to(dev).eval()
    with torch.no_grad():
        cv = cat.sweep(t(feats1, dev), t(grids1, dev), None)
    assert cv.shape == (2, 48, 70, 1, 1) and relmax(cv.cpu().numpy(), orc.sweep_concat(feats1, grids1)) <= TOL
    gm1 = gen.uniform(0, 1, (2, 3, 70, 1, 1)) > 0.3
    masks1 = np.ones((2, 3, 1, 2, 3), np.float32)
    std = modules.SphericalSweepStdMasked(3, 16, 3).to(dev).eval()
    with torch.no_grad():
        sv = std.sweep(t(feats1, dev), t(grids1, dev), t(gm1[..., None], dev), t(masks1, dev))
    ref = orc.sweep_std(feats1, grids1, gm1, masks1)
    assert np.array_equal(sv.cpu().numpy() == 0, ref == 0) and relmax(sv.cpu().numpy(), ref) <= TOL


# ------------------------------------------------------------------ 3. full-size properties (BASELINE.json shapes)
@pytest.mark.parametrize("name,B", [("G16V", 4), ("4CAM", 2), ("HIGHD", 1)])
def test_full_size_std_properties(dev, name, B):
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.WORKLOADS[name]
    rig = w.rig()
    gen = torch.Generator(device=dev).manual_seed(1234)
    feats = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
    masks = t(synthetic.fisheye_footprint_masks(w)[None], dev)
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    with torch.no_grad():
        vol, valid = m.sweep_with_valid(feats, None, None, masks)
        vol2 = m.sweep(feats * 2.0, None, None, masks)          # variance is homogeneous of degree 2: x4, exactly
        perm = torch.randperm(B, device=dev)
        vol_p = m.sweep(feats[perm], None, None, masks)          # frames are independent
    assert vol.shape == (B, w.feat_chs, w.num_cands, *w.cv_hw)
    assert torch.isfinite(vol).all() and (vol >= 0).all()
    # zero pattern == validity mask (a single channel of a valid voxel may round to exactly 0: 1 in 1e8 at HIGHD)
    assert torch.equal((vol != 0).any(dim=1), valid), "zero pattern == validity mask"
    assert ((vol != 0) != valid[:, None].expand_as(vol)).sum().item() <= vol.numel() // 10_000_000
    assert 0.2 < valid.float().mean().item() < 0.999
    assert torch.equal(vol2, vol * 4.0)
    assert torch.equal(vol_p, vol[perm])
    # validity is independent of the frame when masks/geometry are shared
    assert torch.equal(valid, valid[:1].expand_as(valid))


def test_full_size_concat_linearity(dev):
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.WORKLOADS["G8"]
    B = 4
    gen = torch.Generator(device=dev).manual_seed(7)
    f1 = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
    f2 = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
    m = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval().attach_rig(w.rig())
    with torch.no_grad():
        v1, v2 = m.sweep(f1, None), m.sweep(f2, None)
        v12 = m.sweep(f1 * 2.0 + f2, None)
    assert v1.shape == (B, 48, 8, 80, 320)
    assert (v12 - (2.0 * v1 + v2)).abs().max().item() <= 1e-5 * v12.abs().max().item()
    assert torch.equal(m.sweep(f1 * 0.5, None), v1 * 0.5)


def test_full_size_softargmin_properties(dev):
    from mvs_gi_b200 import calibration as cal
    from mvs_gi_b200 import modules

    B, D = 8, 16
    gen = torch.Generator(device=dev).manual_seed(99)
    costs = torch.randn((B, 1, D, 80, 320), device=dev, generator=gen) * 3
    reg = modules.DistanceRegressorWithFixedCandidates(96, cal.DIST_16CPD, 2, True).to(dev)
    with torch.no_grad():
        inv, nc, am = reg.regress(costs, want_norm_costs=True, want_argmax=True)
        up = torch.nn.functional.interpolate(costs[:, 0].cpu(), scale_factor=2, mode="bilinear")
    assert inv.shape == (B, 1, 160, 640) and nc.shape == (B, D, 160, 640)
    assert (nc.sum(1) - 1).abs().max().item() <= 1e-5
    assert inv.min().item() >= reg.inv_dist_idx_min * (1 - 1e-6) and inv.max().item() <= reg.inv_dist_idx_max * (1 + 1e-6)
    assert torch.equal(am.cpu().long(), up.argmax(1)), "arg-max candidate vs PyTorch CPU upsample: bit-exact"
    assert torch.equal(nc.argmax(1).cpu(), up.argmax(1))
    ref_inv = (torch.softmax(up, 1) * reg.inv_dist_idx.cpu()).sum(1, keepdim=True)
    assert (inv.cpu() - ref_inv).abs().max().item() <= TOL * ref_inv.abs().max().item()


# ------------------------------------------------------------------ 4. backward
# fp32 atomics accumulate in a non-deterministic order, so gradients are held to 1e-5 * max|ref| (not bit-exact);
# "exactly zero where nothing samples" IS exact.
@pytest.mark.parametrize("name", ["std3", "std4"])
def test_golden_sweep_std_backward(golden, dev, name):
    from mvs_gi_b200 import modules

    g, gb = golden(f"sweep_{name}"), golden(f"sweep_{name}_bwd")
    B, CAM, C = g["feats"].shape[:3]
    m = modules.SphericalSweepStdMasked(CAM, C, 3).to(dev).eval()
    feats = t(g["feats"], dev).requires_grad_(True)
    vol = m.sweep(feats, t(g["grids"], dev), t(g["grid_masks"], dev), t(g["masks"], dev))
    vol.backward(t(gb["grad_vol"], dev))
    gf, ref = feats.grad.cpu().numpy(), gb["grad_feats"]
    assert np.array_equal(gf == 0, ref == 0)
    assert relmax(gf, ref) <= TOL


def test_golden_sweep_concat_backward(golden, dev):
    from mvs_gi_b200 import modules

    g, gb = golden("sweep_concat3"), golden("sweep_concat3_bwd")
    B, CAM, C = g["feats"].shape[:3]
    m = modules.SphericalSweep(CAM, CAM * C, 3).to(dev).eval()
    feats = t(g["feats"], dev).requires_grad_(True)
    m.sweep(feats, t(g["grids"], dev), None).backward(t(gb["grad_vol"], dev))
    assert relmax(feats.grad.cpu().numpy(), gb["grad_feats"]) <= TOL


@pytest.mark.parametrize("name", ["sa_d16_s2", "sa_d8_s0", "sa_d8_s2_nopre", "sa_d5_s3"])
def test_golden_softargmin_backward(golden, dev, name):
    from mvs_gi_b200 import modules

    g, gb = golden(f"softargmin_{name}"), golden(f"softargmin_{name}_bwd")
    reg = modules.DistanceRegressorWithFixedCandidates(96, list(g["dist"]), float(g["scale"]),
                                                       bool(g["pre_interp"])).to(dev)
    costs = t(g["costs"], dev).requires_grad_(True)
    inv, nc = reg(costs)
    (both,) = torch.autograd.grad([inv, nc], costs, [t(gb["g_inv"], dev), t(gb["g_nc"], dev)], retain_graph=True)
    (only,) = torch.autograd.grad(inv, costs, t(gb["g_inv"], dev))
    assert relmax(both.cpu().numpy(), gb["grad_costs"]) <= TOL
    assert relmax(only.cpu().numpy(), gb["grad_costs_inv_only"]) <= TOL


@pytest.mark.parametrize("name,kw,layout", [
    ("G16V", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), "nchw"),
    ("4CAM", dict(feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=16), "nhwc"),
    ("G16VV_C48", dict(feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=6), "nchw"),
    ("G16V", dict(feat_hw=(17, 70), mask_hw=(40, 130), cv_hw=(9, 37), n_cands=3), "nchw"),
])
def test_sweep_backward_vs_oracle_fused_projection(dev, name, kw, layout):
    """In-kernel projection + per-frame masks; the oracle consumes the kernel's own grids (as in the forward tests)."""
    from mvs_gi_b200 import modules, ops
    from oracle import oracle as orc

    B = 2
    w, feats, _, masks = _case(name, B, 23, **kw)
    rig = w.rig()
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    f = t(feats, dev)
    if layout == "nhwc":
        f = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    f.requires_grad_(True)
    vol = m.sweep(f, None, None, t(masks, dev))
    gv = np.random.default_rng(5).standard_normal(tuple(vol.shape)).astype(np.float32)
    vol.backward(t(gv, dev))
    grids, gm = ops.project_only(rig.to_struct(), dev)
    ref = orc.sweep_std_backward(feats, grids.cpu().numpy()[None], gm.cpu().numpy()[None], masks, gv)
    gf = f.grad.cpu().numpy()
    assert np.array_equal(gf == 0, ref == 0)
    assert relmax(gf, ref) <= TOL


def test_sweep_backward_properties_full_size(dev):
    """G16V at full size (B=2): Euler identity for the degree-2 std volume, and linearity of the backward in grad_vol."""
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.WORKLOADS["G16V"]
    rig = w.rig()
    B = 2
    feats = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, requires_grad=True)
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None]
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    vol = m.sweep(feats, None, None, masks)
    g1, g2 = torch.randn_like(vol), torch.randn_like(vol)
    (a,) = torch.autograd.grad(vol, feats, g1, retain_graph=True)
    (b,) = torch.autograd.grad(vol, feats, g2, retain_graph=True)
    (c,) = torch.autograd.grad(vol, feats, g1 + 2 * g2)
    assert float((c - (a + 2 * b)).abs().max()) <= 1e-4 * float(c.abs().max())
    lhs = float((feats.detach().double() * a.double()).sum())
    rhs = 2.0 * float((vol.detach().double() * g1.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * abs(rhs)


def test_concat_backward_adjoint_full_size(dev):
    """G8 concat at full size: <sweep(f), g> == <f, sweep^T(g)> (the concat sweep is linear)."""
    from mvs_gi_b200 import modules, synthetic

    w = synthetic.WORKLOADS["G8"]
    rig = w.rig()
    feats = torch.randn((2, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, requires_grad=True)
    m = modules.SphericalSweep(w.num_cams, w.num_cams * w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    vol = m.sweep(feats, None)
    g = torch.randn_like(vol)
    (gf,) = torch.autograd.grad(vol, feats, g)
    lhs = float((vol.detach().double() * g.double()).sum())
    rhs = float((feats.detach().double() * gf.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)


def test_softargmin_backward_vs_oracle_and_torch(dev):
    """D=16 x2 at a ragged size vs the oracle, and against torch autograd of the reference formulation on the GPU."""
    import torch.nn.functional as F

    from mvs_gi_b200 import modules, synthetic
    from oracle import oracle as orc

    w = synthetic.WORKLOADS["G16V"]
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, 2, True).to(dev)
    costs_np = (np.random.default_rng(3).standard_normal((2, 2, 16, 13, 37)) * 2).astype(np.float32)  # 2 channels: only [:,0] is used
    costs = t(costs_np, dev).requires_grad_(True)
    inv, nc = reg(costs)
    g_inv, g_nc = torch.randn_like(inv), torch.randn_like(nc)
    (g,) = torch.autograd.grad([inv, nc], costs, [g_inv, g_nc])
    assert float(g[:, 1].abs().max()) == 0.0
    ref = orc.softargmin_backward(costs_np[:, :1], reg.inv_dist_idx.cpu().numpy(), 2, True, g_inv.cpu().numpy(),
                                  g_nc.cpu().numpy())
    assert relmax(g[:, :1].cpu().numpy(), ref) <= TOL
    c2 = t(costs_np, dev).requires_grad_(True)
    up = F.interpolate(c2[:, 0], scale_factor=2, mode="bilinear")
    p = F.softmax(up, 1)
    inv_t = torch.sum(p * reg.inv_dist_idx.expand_as(p), 1, keepdim=True)
    (gt,) = torch.autograd.grad([inv_t, p], c2, [g_inv, g_nc])
    assert relmax(g.cpu().numpy(), gt.cpu().numpy()) <= TOL
    # more candidates than one shared-memory chunk (40 > 16), x3 upsample, sizes that divide no tile, batch 3,
    # and only one of the two upstream gradients present
    from mvs_gi_b200 import calibration as cal
    for D, scale, hw in [(40, 3, (7, 71)), (64, 2, (5, 33)), (3, 2, (1, 1))]:
        reg2 = modules.DistanceRegressorWithFixedCandidates(96, cal.dist_inverse_spaced(D), scale, True).to(dev)
        cn = (np.random.default_rng(D).standard_normal((3, 1, D, *hw)) * 3).astype(np.float32)
        c3 = t(cn, dev).requires_grad_(True)
        inv3, nc3 = reg2(c3)
        gi3, gn3 = torch.randn_like(inv3), torch.randn_like(nc3)
        (g3,) = torch.autograd.grad([inv3, nc3], c3, [gi3, gn3], retain_graph=True)
        ref3 = orc.softargmin_backward(cn, reg2.inv_dist_idx.cpu().numpy(), scale, True, gi3.cpu().numpy(),
                                       gn3.cpu().numpy())
        assert relmax(g3.cpu().numpy(), ref3) <= TOL, (D, scale)
        (g4,) = torch.autograd.grad([inv3], c3, [gi3])
        ref4 = orc.softargmin_backward(cn, reg2.inv_dist_idx.cpu().numpy(), scale, True, gi3.cpu().numpy(), None)
        assert relmax(g4.cpu().numpy(), ref4) <= TOL, (D, scale, "inv only")


@pytest.mark.parametrize("B,D,hw,with_nc", [(2, 16, (80, 320), False), (2, 16, (80, 320), True), (3, 8, (20, 80), True),
                                            (2, 11, (9, 63), True), (1, 5, (4, 31), False), (2, 16, (5, 32), True),
                                            (1, 16, (1, 7), True), (1, 16, (6, 1), True)])
def test_softargmin_backward_gather_kernel_vs_oracle(dev, B, D, hw, with_nc):
    """The x2 gather kernel (csrc/softargmin_bwd.cuh: no atomics, every input pixel written once) at the full G16V size
    and at sizes that put tile seams, the clamped last row / column and the lone first row / column everywhere
    (distance_regressor.py:52-78 through autograd in the reference)."""
    from mvs_gi_b200 import calibration as cal, modules
    from oracle import oracle as orc

    reg = modules.DistanceRegressorWithFixedCandidates(96, cal.dist_inverse_spaced(D), 2, True).to(dev)
    rng = np.random.default_rng(17 * D + hw[0])
    cn = (rng.standard_normal((B, 1, D, *hw)) * 2.5).astype(np.float32)
    c = t(cn, dev).requires_grad_(True)
    inv, nc = reg(c)
    gi, gn = torch.randn_like(inv), torch.randn_like(nc)
    (g,) = torch.autograd.grad([inv, nc] if with_nc else [inv], c, [gi, gn] if with_nc else [gi])
    ref = orc.softargmin_backward(cn, reg.inv_dist_idx.cpu().numpy(), 2, True, gi.cpu().numpy(),
                                  gn.cpu().numpy() if with_nc else None)
    assert relmax(g.cpu().numpy(), ref) <= TOL
    # a gradient buffer that is not zero beforehand must not leak into the result: every element is overwritten
    (g2,) = torch.autograd.grad([reg(c)[0]], c, [gi])
    ref2 = orc.softargmin_backward(cn, reg.inv_dist_idx.cpu().numpy(), 2, True, gi.cpu().numpy(), None)
    assert relmax(g2.cpu().numpy(), ref2) <= TOL


def test_bf16_volume_stated_tolerance(dev):
    """MVSGI_FLAG_VOL_BF16 (row f2): the bfloat16 channels-last volume is the fp32 volume rounded to nearest even —
    bit for bit (so |vol_bf16 - vol| <= 2^-8 |vol| and the zero pattern is unchanged) — for NCHW-staged and in-place
    features; post_vol in bfloat16 stays within 2e-2 * max of the fp32 module."""
    from mvs_gi_b200 import modules, ops, synthetic

    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=8)
    B = 2
    feats = t(synthetic.make_feats(w, B, 51), dev)
    masks = t(synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, 52), dev)
    ref = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(w.rig())
    ref.set_vol_channels_last(True)
    m16 = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(w.rig())
    m16.post_vol.load_state_dict(ref.post_vol.state_dict())
    m16.set_vol_bf16(True)
    with torch.no_grad():
        v32 = ref.sweep(feats, None, None, masks)
        for f in (feats, feats.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)):
            v16 = m16.sweep(f, None, None, masks)
            assert v16.dtype == torch.bfloat16 and v16.is_contiguous(memory_format=torch.channels_last_3d)
            assert torch.equal(v16, v32.to(torch.bfloat16)), "bf16 volume == round-to-nearest-even of the fp32 volume"
        o32, o16 = ref(feats, None, None, masks), m16(feats, None, None, masks)
    assert o16.dtype == torch.float32
    assert float((o16 - o32).abs().max()) <= 2e-2 * float(o32.abs().max())
    with pytest.raises(RuntimeError):
        ops.SweepCall(feats, None, None, masks, ops.rig_to_blob(w.rig().to_struct()), 0, *w.cv_hw, 8, vol_bf16=True)
