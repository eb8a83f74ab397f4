"""Every BASELINE.json configuration at its FULL size, CUDA path against the C oracle (oracle/mvsgi_oracle.c).

The oracle is fed the kernel's own grids (mvsgi_project_only_f32; the grids themselves are bounded against float64 in
test_gpu_parity.py::test_projection_accuracy), so the SURVEY.md §8(c) criteria apply unchanged at full size:
validity mask / zero pattern bit-exact, max|vol - ref| <= 1e-5 * max|ref|, arg-max candidate bit-exact,
max|inv_dist - ref| <= 1e-5 * max|ref|.  What only full size exercises: 32-bit offset arithmetic near its limits,
all 296 CTAs with interleaved runs, the 80 KB tap table, candidate chunking of the high-D volume.
The OpenMP oracle needs about a second per G16V frame on the box's host cores; batches are 1-2 frames."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvs_gi_b200 import abi

    abi.load_library()
    return torch.device("cuda:0")


def relmax(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _nhwc(f):
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


# name, frames, features handed over channels-last (tapped in place), channels-last volume
STD_CASES = [
    ("G16V", 2, False, False),       # config29, the reference's layouts: NCHW features staged, NCDHW volume
    ("G16V", 2, True, True),         # the bench default: in-place taps, channels-last volume
    ("G16VV_C48", 1, True, True),    # config103 read literally: 48-channel std
    ("G16VV_C48", 1, False, False),
    ("4CAM", 2, True, True),         # map_cam_frame_4cam_with_surrogate_input
    ("4CAM", 1, False, False),
    ("G16V_DS", 2, True, False),     # raw fisheyes, DoubleSphere projected in-kernel
    ("HIGHD", 1, True, True),        # 64 candidates at 2x resolution (419 MB volume)
    ("HIGHD", 1, False, False),
]


@pytest.mark.parametrize("name,B,nhwc,vol_cl", STD_CASES)
def test_std_sweep_full_size_vs_oracle(dev, name, B, nhwc, vol_cl):
    from mvs_gi_b200 import modules, ops, synthetic
    from oracle import oracle as orc

    w = synthetic.WORKLOADS[name]
    feats = synthetic.make_feats(w, B, seed=101)
    masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, seed=102)
    rig = w.rig()
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    m.set_vol_channels_last(vol_cl)
    f = torch.from_numpy(feats).to(dev)
    with torch.no_grad():
        vol, valid = m.sweep_with_valid(_nhwc(f) if nhwc else f, None, None, torch.from_numpy(masks).to(dev))
        grids, gm = ops.project_only(rig.to_struct(), dev)
    ref, ref_valid = orc.sweep_std(feats, grids.cpu().numpy()[None], gm.cpu().numpy()[None], masks, return_valid=True)
    vol = vol.cpu().numpy()
    assert np.array_equal(valid.cpu().numpy(), ref_valid), "validity mask must be bit-exact"
    assert np.array_equal(vol == 0, ref == 0), "zero pattern must be bit-exact"
    assert 0.02 < ref_valid.mean() < 0.999
    assert relmax(vol, ref) <= TOL


@pytest.mark.parametrize("name,nhwc", [("G8", True), ("G8", False), ("E8", True)])
def test_concat_sweep_full_size_vs_oracle(dev, name, nhwc):
    """E8 / G8 (config24 / config25): concat builder, 8 candidates, 48 output channels.  The sampler itself is
    bit-exact against the oracle (same operation order), so the bound here is 0 up to the fma placement."""
    from mvs_gi_b200 import modules, ops, synthetic
    from oracle import oracle as orc

    w = synthetic.WORKLOADS[name]
    B = 2
    feats = synthetic.make_feats(w, B, seed=103)
    rig = w.rig()
    m = modules.SphericalSweep(w.num_cams, w.out_chs, 3).to(dev).eval().attach_rig(rig)
    f = torch.from_numpy(feats).to(dev)
    with torch.no_grad():
        vol = m.sweep(_nhwc(f) if nhwc else f, None, None).cpu().numpy()
        grids, _ = ops.project_only(rig.to_struct(), dev)
    ref = orc.sweep_concat(feats, grids.cpu().numpy()[None])
    assert vol.shape == ref.shape == (B, 48, 8, 80, 320)
    assert np.array_equal(vol == 0, ref == 0)
    assert relmax(vol, ref) <= TOL


@pytest.mark.parametrize("name,B", [("G16V", 2), ("G8", 2), ("HIGHD", 1)])
def test_softargmin_full_size_vs_oracle(dev, name, B):
    """DistanceRegressorWithFixedCandidates at the configuration's full size (x2 pre-interpolation): arg-max candidate
    bit-exact, inv_dist and norm_costs within 1e-5."""
    from mvs_gi_b200 import modules, synthetic
    from oracle import oracle as orc

    w = synthetic.WORKLOADS[name]
    costs = synthetic.make_costs(w, B, seed=104) * 3.0
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
    with torch.no_grad():
        inv, nc, am = reg.regress(torch.from_numpy(costs).to(dev), want_norm_costs=True, want_argmax=True)
    r_inv, r_nc, r_am = orc.softargmin(costs, reg.inv_dist_idx.cpu().numpy(), w.scale, True)
    assert np.array_equal(am.cpu().numpy(), r_am), "arg-max candidate must be bit-exact"
    assert relmax(inv.cpu().numpy(), r_inv) <= TOL
    assert np.abs(nc.cpu().numpy() - r_nc).max() <= TOL


def test_linear_sphere_rig_is_swept(dev):
    """A rig of LinearSphere fisheyes (camera_models.py:1489-1548) through the fused sweep: the third camera model the
    in-kernel projection supports was only exercised by the resampler before."""
    from mvs_gi_b200 import calibration as cal
    from mvs_gi_b200 import modules, ops, synthetic
    from oracle import oracle as orc

    base = synthetic.WORKLOADS["G16V"]
    cams = [cal.LinearSphereCamera(195.0, 608, 608) for _ in range(3)]
    rig = cal.Rig(cal.EquirectCamera(40, 160), cams, cal.fixture_poses(3, same_facing=False), base.dist_list[::2])
    B, C, Hf, Wf = 2, 16, 152, 152
    gen = np.random.default_rng(7)
    feats = gen.standard_normal((B, 3, C, Hf, Wf), dtype=np.float32)
    yy, xx = np.mgrid[0:608, 0:608].astype(np.float32)
    disc = (np.hypot(xx + 0.5 - 304, yy + 0.5 - 304) <= 0.49 * 608).astype(np.float32)
    masks = synthetic.erase_rectangles(np.broadcast_to(disc, (3, 1, 608, 608)).copy(), B, seed=8)
    m = modules.SphericalSweepStdMasked(3, C, 3).to(dev).eval().attach_rig(rig)
    f = torch.from_numpy(feats).to(dev)
    with torch.no_grad():
        vol, valid = m.sweep_with_valid(f, None, None, torch.from_numpy(masks).to(dev))
        vol2 = m.sweep(_nhwc(f), None, None, torch.from_numpy(masks).to(dev))
        grids, gm = ops.project_only(rig.to_struct(), dev)
    # the kernel's LinearSphere grids against the oracle's fp32 / float64 evaluation of camera_models.py:1506-1548
    g32, m32 = orc.make_grids(rig.to_struct())
    g64, m64 = orc.make_grids_f64(rig.to_struct())
    gk, mk = grids.cpu().numpy(), gm.cpu().numpy()[..., 0]
    assert np.abs(gk - g64)[mk & m64].max() <= 2e-5
    assert int((mk != m32).sum()) <= max(2, mk.size // 5000)
    ref, ref_valid = orc.sweep_std(feats, gk[None], gm.cpu().numpy()[None], masks, return_valid=True)
    vol = vol.cpu().numpy()
    assert np.array_equal(valid.cpu().numpy(), ref_valid)
    assert np.array_equal(vol == 0, ref == 0)
    assert 0.02 < ref_valid.mean() < 0.999
    assert relmax(vol, ref) <= TOL
    assert torch.equal(vol2.cpu(), torch.from_numpy(vol))
