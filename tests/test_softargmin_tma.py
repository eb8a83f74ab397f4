"""The TMA-staged scale-factor-2 soft-argmin (csrc/softargmin_tma.cuh) against the pointer-per-load kernel
(csrc/softargmin.cuh, the one the golden-vector tests of tests/test_gpu_parity.py were written for).

The library picks the TMA kernel when the cost volume can be described by a tensor map (unit column stride, 16-byte
aligned base and pitches).  A view whose base is 4 bytes off that alignment makes it fall back to the pointer-per-load
kernel on the same values, so the two outputs can be compared bit for bit — inv_dist, norm_costs and the arg-max
(distance_regressor.py:52-78).  Shapes cover the border blocks (first / last output row and column on their own),
candidate counts below and at the register-tile sizes (8, 16), and sizes that divide nothing, and batches
with several tiles per persistent CTA (both ring stages reused, both mbarrier parities)."""
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


def _misaligned_copy(costs):
    """Same values, contiguous strides, base address 4 bytes past a 16-byte boundary: not describable by a tensor map."""
    flat = torch.empty(costs.numel() + 1, dtype=costs.dtype, device=costs.device)
    v = flat[1:].view(costs.shape)
    v.copy_(costs)
    assert v.data_ptr() % 16 == 4
    return v


@pytest.mark.parametrize("B,D,H,W", [(2, 16, 80, 320), (1, 16, 7, 12), (3, 8, 20, 80), (2, 5, 9, 36), (1, 13, 33, 68),
                                     (2, 16, 4, 32), (1, 16, 5, 4),
                                     # more tiles than resident CTAs: every CTA walks several tiles through its ring
                                     (12, 16, 80, 320), (9, 11, 80, 320), (64, 8, 40, 100)])
@pytest.mark.parametrize("want_nc,want_am", [(False, False), (True, True)])
def test_tma_kernel_matches_pointer_kernel_bitwise(dev, B, D, H, W, want_nc, want_am):
    from mvs_gi_b200 import ops

    g = torch.Generator(device=dev).manual_seed(100 * D + H)
    costs = torch.randn((B, 1, D, H, W), device=dev, generator=g) * 3.0
    assert costs.data_ptr() % 16 == 0
    idx = torch.linspace(0.01, 1.0, D, device=dev)
    inv_t, nc_t, am_t = ops.softargmin_op(costs, idx, 2 * H, 2 * W, 0.5, want_nc, want_am)
    inv_p, nc_p, am_p = ops.softargmin_op(_misaligned_copy(costs), idx, 2 * H, 2 * W, 0.5, want_nc, want_am)
    torch.cuda.synchronize()
    assert torch.equal(inv_t, inv_p), f"inv_dist: max |diff| {float((inv_t - inv_p).abs().max())}"
    if want_nc:
        assert torch.equal(nc_t, nc_p)
        assert torch.allclose(nc_t.sum(1), torch.ones_like(nc_t.sum(1)), atol=1e-5)
    if want_am:
        assert torch.equal(am_t, am_p)


def test_tma_kernel_against_the_oracle(dev):
    """Full G16V shape through the TMA kernel vs the C oracle: arg-max bit-exact, inv_dist within 1e-5 of the maximum."""
    from mvs_gi_b200 import ops
    from oracle import oracle as orc

    B, D, H, W = 2, 16, 80, 320
    rng = np.random.default_rng(5)
    costs = (rng.standard_normal((B, 1, D, H, W)) * 2.0).astype(np.float32)
    idx = np.linspace(0.02, 1.0, D).astype(np.float32)
    inv, nc, am = ops.softargmin_op(torch.from_numpy(costs).to(dev), torch.from_numpy(idx).to(dev), 2 * H, 2 * W, 0.5,
                                    True, True)
    ref_inv, ref_nc, ref_am = orc.softargmin(costs, idx, 2, True)
    assert np.array_equal(am.cpu().numpy(), ref_am)
    assert np.abs(inv.cpu().numpy() - ref_inv).max() <= 1e-5 * np.abs(ref_inv).max()
    assert np.abs(nc.cpu().numpy() - ref_nc).max() <= 1e-5
