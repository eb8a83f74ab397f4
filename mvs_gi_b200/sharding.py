"""Multi-GPU partitioning of the sweep path: one process per GPU, `torch.distributed` for the plumbing.

Frames are independent everywhere on this path (nothing in spherical_sweep_avg.py:38-136, spherical_sweep.py:38-68 or
distance_regressor.py:51-80 crosses the batch dimension), so the natural partition is by FRAME and needs no
collective: `frame_shard` gives each rank a contiguous, balanced range of the global batch.

The optional CANDIDATE-sharded variant (BASELINE.json config 5, "high-D stress") splits the D distance candidates
instead: rank r sweeps candidates [d0, d1) for all frames and the [B, C, D, Ho, Wo] volume the (un-sharded) regulator
needs is re-assembled on every rank with one all-gather over NVLink.  It is NVLink-bound by construction (each rank
receives (G-1)/G of the volume at <= 0.9 TB/s while a single GPU writes the whole volume at ~6.5 TB/s), so it only
makes sense when one GPU cannot hold the volume; DESIGN.md states the measured numbers.

The reference has no equivalent (its only parallelism is Lightning DDP over frames, configs/base_trainer.yaml:2-4).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .calibration import Rig


def balanced_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the `rank`-th of `world` contiguous, balanced pieces of range(n): sizes differ by at most one
    and the larger pieces come first.  Ranks beyond n get an empty range."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n < 0:
        raise ValueError("n must be >= 0")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frame_shard(num_frames: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Frames [start, stop) of the global batch this rank sweeps.  No collective follows: every rank's outputs are
    complete for its own frames."""
    rank, world = _rank_world(rank, world)
    return balanced_range(num_frames, rank, world)


def candidate_shard(num_cands: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Candidates [d0, d1) this rank sweeps in the candidate-sharded variant."""
    rank, world = _rank_world(rank, world)
    return balanced_range(num_cands, rank, world)


def _rank_world(rank, world, group=None):
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
        return 0, 1
    return rank, world


def sub_rig(rig: Rig, d0: int, d1: int) -> Rig:
    """The same rig restricted to candidates [d0, d1): what a rank of the candidate-sharded variant projects with."""
    if not 0 <= d0 < d1 <= rig.num_cands:
        raise ValueError(f"empty or out-of-range candidate slice [{d0}, {d1}) of {rig.num_cands}")
    return Rig(rig.cv_camera, rig.cameras, rig.poses_cv_cam, list(rig.dist_list[d0:d1]))


def all_gather_candidates(local: Tensor, num_cands: int, group=None, out: Optional[Tensor] = None) -> Tensor:
    """Re-assemble the full volume from per-rank candidate slices.

    local : [B, C, Dl, Ho, Wo], this rank's slice = candidates candidate_shard(num_cands, rank, world) in order
            (Dl may differ by one between ranks when world does not divide num_cands).
    ->      [B, C, num_cands, Ho, Wo] on every rank, bit-identical to an un-sharded sweep.

    One `all_gather_into_tensor` into a rank-major staging buffer (slices padded to the widest rank so that the
    collective is a single equal-sized NCCL all-gather) and one strided copy into the [B, C, D, Ho, Wo] layout the
    regulator's Conv3d wants."""
    rank, world = _rank_world(None, None, group)
    B, C, Dl, Ho, Wo = local.shape
    d0, d1 = balanced_range(num_cands, rank, world)
    if Dl != d1 - d0:
        raise RuntimeError(f"rank {rank} holds {Dl} candidates, its shard of {num_cands} over {world} is {d1 - d0}")
    if out is None:
        out = torch.empty((B, C, num_cands, Ho, Wo), dtype=local.dtype, device=local.device)
    elif tuple(out.shape) != (B, C, num_cands, Ho, Wo):
        raise RuntimeError("`out` has the wrong shape")
    if world == 1:
        out.copy_(local)
        return out
    widest = balanced_range(num_cands, 0, world)
    dmax = widest[1] - widest[0]
    if Dl == dmax and local.is_contiguous():
        send = local
    else:
        send = torch.zeros((B, C, dmax, Ho, Wo), dtype=local.dtype, device=local.device)
        send[:, :, :Dl].copy_(local)
    # rank-major staging, declared as the dim-0 concatenation of the slices (the form every backend accepts)
    flat = torch.empty((world * B, C, dmax, Ho, Wo), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, send, group=group)
    staging = flat.view(world, B, C, dmax, Ho, Wo)
    for r in range(world):
        r0, r1 = balanced_range(num_cands, r, world)
        if r1 > r0:
            out[:, :, r0:r1].copy_(staging[r, :, :, :r1 - r0])
    return out


class CandidateShardedSweep(torch.nn.Module):
    """Candidate-sharded front-end of a std / concat sweep module (modules.SphericalSweepStdMasked / SphericalSweep).

    Every rank gets the full batch of features; rank r evaluates only its slice of the candidate list (fused
    in-kernel projection from a sub-rig, or a slice of the supplied grids) and `all_gather_candidates` rebuilds
    the volume.  `post_vol` runs afterwards on the full volume, as in the reference forward()."""

    def __init__(self, builder: torch.nn.Module, rig: Optional[Rig] = None, group=None):
        super().__init__()
        self.builder = builder
        self.group = group
        self.full_rig = rig
        self.num_cands = rig.num_cands if rig is not None else None
        if rig is not None:
            rank, world = _rank_world(None, None, group)
            d0, d1 = balanced_range(rig.num_cands, rank, world)
            if d1 <= d0:
                raise RuntimeError(f"rank {rank} of {world} would get no candidates out of {rig.num_cands}")
            builder.attach_rig(sub_rig(rig, d0, d1))

    def sweep(self, feats: Tensor, grids: Optional[Tensor], grid_masks: Optional[Tensor],
              masks: Optional[Tensor]) -> Tensor:
        rank, world = _rank_world(None, None, self.group)
        if self.full_rig is not None:
            D = self.num_cands
            g = gm = None
        else:
            D = grids.shape[2]
            d0, d1 = balanced_range(D, rank, world)
            g = grids[:, :, d0:d1]
            gm = grid_masks[:, :, d0:d1] if grid_masks is not None else None
        if hasattr(self.builder, "sweep_with_valid"):
            local = self.builder.sweep(feats, g, gm, masks)
        else:
            local = self.builder.sweep(feats, g, masks)
        return all_gather_candidates(local, D, self.group)

    def forward(self, feats, grids, grid_masks, masks):
        return self.builder.post_vol(self.sweep(feats, grids, grid_masks, masks))


def shard_lists(n: int, world: int) -> List[Tuple[int, int]]:
    """All ranks' ranges, for logging and tests."""
    return [balanced_range(n, r, world) for r in range(world)]
