"""The drop-in boundary: constructor signatures, attributes, state_dict keys and `patch_model` against
the reference's own classes when /root/reference is present (build container), against the recorded contract
otherwise (GPU box).  The GPU case runs a whole SphericalSweepStereoBase forward through patched modules."""
import importlib.util
import inspect
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from mvs_gi_b200 import modules, synthetic

REF = Path("/root/reference")

# the contract the reference classes define (checked against them below whenever the reference is importable)
SWEEP_CTOR = [("num_cams", inspect.Parameter.empty), ("feat_chs", inspect.Parameter.empty),
              ("post_k_sz", inspect.Parameter.empty), ("norm_type", "batch"), ("relu_type", "leaky")]
REG_CTOR = [("bf", 96), ("dist_cands", None), ("interp_scale_factor", 0), ("pre_interp", False)]
SWEEP_KEYS = ["post_vol.conv_layer.weight", "post_vol.norm_layer.weight", "post_vol.norm_layer.bias",
              "post_vol.norm_layer.running_mean", "post_vol.norm_layer.running_var",
              "post_vol.norm_layer.num_batches_tracked"]


def ctor(cls):
    return [(n, p.default) for n, p in list(inspect.signature(cls.__init__).parameters.items())[1:]]


def test_module_contract():
    for cls in (modules.SphericalSweepStdMasked, modules.SphericalSweep):
        assert ctor(cls) == SWEEP_CTOR
        m = cls(3, 16, 3)
        assert list(m.state_dict().keys()) == SWEEP_KEYS
        for attr in ("num_cams", "feat_chs", "norm_type", "relu_type", "grid_sample_mode", "post_vol"):
            assert hasattr(m, attr)
        assert m.grid_sample_mode == "bilinear" and m.post_vol.conv_layer.bias is None
    assert list(inspect.signature(modules.SphericalSweepStdMasked.sweep).parameters)[1:] == \
        ["feats", "grids", "grid_masks", "masks"]
    assert list(inspect.signature(modules.SphericalSweep.sweep).parameters)[1:] == ["feats", "grids", "masks"]
    for cls in (modules.SphericalSweepStdMasked, modules.SphericalSweep):
        assert list(inspect.signature(cls.forward).parameters)[1:] == ["feats", "grids", "grid_masks", "masks"]
    got = ctor(modules.DistanceRegressorWithFixedCandidates)
    assert [n for n, _ in got] == [n for n, _ in REG_CTOR]
    assert got[0][1] == 96 and got[2][1] == 0 and got[3][1] is False
    r = modules.DistanceRegressorWithFixedCandidates(96, [0.5, 1, 2, 100], 2, True)
    assert list(r.state_dict().keys()) == ["inv_dist_idx"] and tuple(r.inv_dist_idx.shape) == (1, 4, 1, 1)
    assert r.inv_dist_idx_min == pytest.approx(0.96) and r.inv_dist_idx_max == pytest.approx(192.0)
    r.update_dist_cands([1, 2, 4, 8])
    assert r.inv_dist_idx.reshape(-1).tolist() == [96.0, 48.0, 24.0, 12.0] and r.inv_dist_idx_max == 96.0
    # there is no CPU path: CPU tensors are an error, not a fallback
    with pytest.raises(RuntimeError, match="CUDA"):
        r(torch.zeros(1, 1, 4, 3, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        modules.SphericalSweep(3, 48, 3).sweep(torch.zeros(1, 3, 16, 4, 8), torch.zeros(1, 3, 2, 3, 5, 2), None)


def _reference():
    if not REF.exists():
        pytest.skip("the reference checkout is only present in the build container")
    sys.path.insert(0, str(REF))
    from dsta_mvs.model.cost_volume_builder import SphericalSweep, SphericalSweepStdMasked
    from dsta_mvs.model.distance_regressor.distance_regressor import DistanceRegressorWithFixedCandidates
    spec = importlib.util.spec_from_file_location("ref_torch_only", REF / "dsta_mvs/model/mvs_model/torch_only.py")
    torch_only = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(torch_only)
    return SphericalSweepStdMasked, SphericalSweep, DistanceRegressorWithFixedCandidates, torch_only


def test_contract_is_the_references_and_patch_model_keeps_checkpoints():
    RStd, RCat, RReg, torch_only = _reference()
    assert ctor(RStd) == SWEEP_CTOR and ctor(RCat) == SWEEP_CTOR
    assert [n for n, _ in ctor(RReg)] == [n for n, _ in REG_CTOR]
    assert list(inspect.signature(RStd.sweep).parameters) == list(inspect.signature(modules.SphericalSweepStdMasked.sweep).parameters)
    assert list(inspect.signature(RCat.sweep).parameters) == list(inspect.signature(modules.SphericalSweep.sweep).parameters)
    assert list(inspect.signature(RStd.forward).parameters) == list(inspect.signature(modules.SphericalSweepStdMasked.forward).parameters)
    torch.manual_seed(0)
    for RB, feat_chs in ((RStd, 16), (RCat, 48)):
        ref_b = RB(3, feat_chs, 3)
        ref_b.post_vol.norm_layer.running_mean.normal_()
        ref_b.post_vol.norm_layer.running_var.uniform_(0.5, 2.0)
        ref_r = RReg(96, synthetic.WORKLOADS["G16V"].dist_list, 2, True)
        host = torch_only.SphericalSweepStereoBase(nn.Identity(), ref_b, nn.Identity(), ref_r)
        before = {k: v.clone() for k, v in host.state_dict().items()}
        host.eval()
        out = modules.patch_model(host, return_norm_costs=False)
        assert out is host
        assert type(host.cv_builder).__module__.startswith("mvs_gi_b200") and type(host.cv_builder).__name__ == RB.__name__
        assert type(host.dist_regressor).__module__.startswith("mvs_gi_b200")
        after = host.state_dict()
        assert list(after.keys()) == list(before.keys())
        for k in before:
            assert torch.equal(after[k], before[k]), k
        assert host.cv_builder.training is False
        host.load_state_dict(before)  # a checkpoint written by the reference loads into the patched model
        assert host.dist_regressor.return_norm_costs is False
        assert host.dist_regressor.bf == ref_r.bf and host.dist_regressor.pre_interp == ref_r.pre_interp
        assert host.dist_regressor.inv_dist_idx_min == ref_r.inv_dist_idx_min
        # the Lightning wrappers hold the model in `.mvs_model` (spherical_sweep_stereo_2024.py:14)
        wrapper = nn.Module()
        wrapper.mvs_model = torch_only.SphericalSweepStereoBase(nn.Identity(), RB(3, feat_chs, 3), nn.Identity(),
                                                                RReg(96, [0.5, 1, 100], 0, False))
        modules.patch_model(wrapper)
        assert type(wrapper.mvs_model.cv_builder).__module__.startswith("mvs_gi_b200")
    with pytest.raises(TypeError):
        modules.patch_model(torch_only.SphericalSweepStereoBase(nn.Identity(), nn.Identity(), nn.Identity(), ref_r))


class _Host(nn.Module):
    """SphericalSweepStereoBase.forward (dsta_mvs/model/mvs_model/torch_only.py:30-35) with stand-in extractor and
    regulator, for the GPU box where the reference is absent."""

    def __init__(self, cv_builder, dist_regressor, feat_chs):
        super().__init__()
        self.feature_extractor = nn.Conv2d(3, feat_chs, 3, padding=1)
        self.cv_builder = cv_builder
        self.cv_regulator = nn.Conv3d(feat_chs, 1, 3, padding=1)
        self.dist_regressor = dist_regressor

    def forward(self, imgs, grids, grid_masks, masks):
        og = imgs.shape[:2]
        feats = self.feature_extractor(imgs.view(og[0] * og[1], *imgs.shape[2:]))
        feats = feats.view(*og, *feats.shape[1:])
        vol = self.cv_builder(feats, grids, grid_masks, masks)
        costs = self.cv_regulator(vol)
        return self.dist_regressor(costs)


@pytest.mark.gpu
def test_whole_model_forward_and_backward_through_the_dropins():
    """imgs -> extractor -> B200 sweep -> post_vol -> regulator -> B200 soft-argmin, eval and train, against the same
    host model holding torch-eager restatements of the two reference modules (oracle/ref_port_torch.py)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import oracle as orc
    from oracle import ref_port_torch as port

    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(32, 128), mask_hw=(32, 128), cv_hw=(20, 80), n_cands=8)
    g, gm = orc.make_grids(w.rig().to_struct())
    B = 2
    grids = torch.from_numpy(g)[None].repeat(B, 1, 1, 1, 1, 1).to(dev)
    gmask = torch.from_numpy(gm)[None, ..., None].repeat(B, 1, 1, 1, 1, 1).to(dev)
    masks = torch.from_numpy(synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, 3)).to(dev)
    torch.manual_seed(1)
    imgs = torch.rand(B, w.num_cams, 3, *w.feat_hw, device=dev)
    builder = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3)
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, 2, True)
    host = _Host(builder, reg, w.feat_chs).to(dev)

    class RefBuilder(nn.Module):
        def __init__(self, post_vol):
            super().__init__()
            self.post_vol = post_vol

        def forward(self, feats, grids, grid_masks, masks):
            return self.post_vol(port.sweep_std(feats, grids, grid_masks, masks))

    class RefReg(nn.Module):
        def forward(self, costs):
            return port.soft_argmin(costs, reg.inv_dist_idx.reshape(-1), 2, True)

    ref_host = _Host(RefBuilder(builder.post_vol), RefReg(), w.feat_chs).to(dev)
    ref_host.feature_extractor, ref_host.cv_regulator = host.feature_extractor, host.cv_regulator
    for layout in (False, True):
        builder.set_vol_channels_last(layout)
        for train in (False, True):
            host.train(train)
            ref_host.train(train)
            for p in host.parameters():
                p.grad = None
            inv, nc = host(imgs, grids, gmask, masks)
            rinv, rnc = ref_host(imgs, grids, gmask, masks)
            assert inv.shape == (B, 1, 40, 160) and nc.shape == (B, w.num_cands, 40, 160)
            assert float((inv - rinv).detach().abs().max()) <= 2e-4 * float(rinv.detach().abs().max())
            assert float((nc - rnc).detach().abs().max()) <= 2e-4
            if train:
                loss = inv.mean() + (nc * nc).mean()
                loss.backward()
                got = [p.grad.clone() for p in host.feature_extractor.parameters()]
                for p in host.parameters():
                    p.grad = None
                rloss = rinv.mean() + (rnc * rnc).mean()
                rloss.backward()
                want = [p.grad for p in host.feature_extractor.parameters()]
                for a, b in zip(got, want):
                    assert float((a - b).abs().max()) <= 2e-3 * float(b.abs().max()) + 1e-7


@pytest.mark.gpu
def test_masks_are_packed_once_per_tensor_and_checked_on_request():
    """The module keeps the packed bits of the masks tensor it last saw (same object, same version): a second forward
    with the same masks launches no packing kernel; an in-place edit or a new tensor re-packs; negative / NaN masks
    raise when `check_masks` is on (and always through set_static_masks)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mvs_gi_b200 import modules, ops, synthetic

    dev = torch.device("cuda:0")
    w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(10, 40), n_cands=4)
    B = 2
    feats = torch.from_numpy(synthetic.make_feats(w, B, seed=2)).to(dev)
    masks = torch.from_numpy(synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, seed=3)).to(dev)
    m = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(w.rig())
    n0 = ops.launch_counter()
    v1 = m.sweep(feats, None, None, masks)
    n1 = ops.launch_counter()
    v2 = m.sweep(feats, None, None, masks)
    n2 = ops.launch_counter()
    assert n1 - n0 == (n2 - n1) + 1, "the second call with the same masks tensor must not pack again"
    assert torch.equal(v1, v2)
    masks[0, 1, 0, :, :] = 0.0  # in-place edit bumps the version -> re-pack, different volume
    v3 = m.sweep(feats, None, None, masks)
    n3 = ops.launch_counter()
    assert n3 - n2 == n1 - n0
    assert not torch.equal(v3, v1)
    ref = ops.sweep_op(feats, None, None, masks, ops.rig_to_blob(w.rig().to_struct()), 0, *w.cv_hw, 4, False)[0]
    assert torch.equal(v3, ref), "cached bits == float masks through the op"
    # expanded (stride-0) batch of one mask set == the set itself
    one = masks[:1].clone()
    assert torch.equal(m.sweep(feats, None, None, one.expand(B, -1, -1, -1, -1)), m.sweep(feats, None, None, one))
    bad = masks.clone()
    bad[0, 0, 0, 3, 5] = -1.0
    m.sweep(feats, None, None, bad)  # unchecked by default (no sync, graph-capturable)
    m.check_masks = True
    bad2 = bad.clone()
    with pytest.raises(RuntimeError, match="negative or NaN"):
        m.sweep(feats, None, None, bad2)
    nan = masks[:1].clone()
    nan[0, 0, 0, 0, 0] = float("nan")
    with pytest.raises(RuntimeError, match="negative or NaN"):
        m.set_static_masks(nan)
