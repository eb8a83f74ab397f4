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


class _RawCudaBuffer:
    """Device memory owned by libmvsgi_b200 (mvsgi_p2p_alloc), exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner  # keeps the allocation alive as long as any tensor view exists


class PeerVolume:
    """`n_buffers` copies of a [B, C, D, Ho, Wo] float32 volume on this GPU, each also mapped into every other rank of
    `group` (one process per GPU on one NVLink/NVSwitch box), so that a candidate-sharded sweep can store its slice
    straight into all ranks' volumes (include/mvsgi_b200.h: MvsgiSweepArgs.peer_vols).

    Buffers alternate between calls: with the per-call rendezvous of CandidateShardedSweep a rank can be at most one
    call ahead of its peers, so two buffers are enough to keep a fast rank from overwriting a volume its neighbour
    is still reading."""

    def __init__(self, shape, device, group=None, n_buffers: int = 2):
        import ctypes as C

        from . import abi

        self.lib = abi.load_library()
        self.group = group
        self.shape = tuple(int(v) for v in shape)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PeerVolume lives in GPU memory; there is no CPU path")
        rank, world = _rank_world(None, None, group)
        self.rank, self.world = rank, world
        if world - 1 > abi.MAX_PEERS:
            raise RuntimeError(f"at most {abi.MAX_PEERS + 1} ranks")
        n = 1
        for v in self.shape:
            n *= v
        self.stride_bytes = (n * 4 + 255) // 256 * 256
        self.n_buffers = n_buffers
        self._turn = 0
        base = C.c_void_p()
        with torch.cuda.device(self.device):
            abi.check(self.lib.mvsgi_p2p_alloc(self.stride_bytes * n_buffers, C.byref(base)), "mvsgi_p2p_alloc")
            self._base = base.value
            handle = (C.c_uint8 * abi.P2P_HANDLE_BYTES)()
            abi.check(self.lib.mvsgi_p2p_export(self._base, handle), "mvsgi_p2p_export")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            everyone = torch.empty(world * abi.P2P_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
            if world > 1:
                dist.all_gather_into_tensor(everyone, mine, group=group)
            else:
                everyone.copy_(mine)
            everyone = everyone.cpu().numpy().reshape(world, abi.P2P_HANDLE_BYTES)
            self._peer_bases = []
            for r in range(world):
                if r == rank:
                    continue
                h = (C.c_uint8 * abi.P2P_HANDLE_BYTES)(*[int(v) for v in everyone[r]])
                ptr = C.c_void_p()
                abi.check(self.lib.mvsgi_p2p_open(h, C.byref(ptr)), f"mvsgi_p2p_open(rank {r})")
                self._peer_bases.append(ptr.value)
        self.tensors = [torch.as_tensor(_RawCudaBuffer(self._base + i * self.stride_bytes, self.shape, self),
                                        device=self.device) for i in range(n_buffers)]
        if world > 1:
            dist.barrier(group=group)  # nobody stores into a peer before everybody has mapped everybody

    def next(self):
        """(local tensor view, peer pointers) of the buffer for this call."""
        i = self._turn
        self._turn = (self._turn + 1) % self.n_buffers
        return self.tensors[i], [b + i * self.stride_bytes for b in self._peer_bases]

    def close(self):
        if getattr(self, "_base", None) is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)  # peers have stopped storing into this allocation
            for p in self._peer_bases:
                self.lib.mvsgi_p2p_close(p)
            self.tensors = []
            self.lib.mvsgi_p2p_free(self._base)
        self._base = None


class CandidateShardedSweep(torch.nn.Module):
    """Candidate-sharded front-end of a std / concat sweep module (modules.SphericalSweepStdMasked / SphericalSweep).

    Every rank gets the full batch of features; rank r evaluates only its slice of the candidate list (fused
    in-kernel projection from a sub-rig, or a slice of the supplied grids) and `all_gather_candidates` rebuilds
    the volume.  `post_vol` runs afterwards on the full volume, as in the reference forward()."""

    def __init__(self, builder: torch.nn.Module, rig: Optional[Rig] = None, group=None, fused: bool = False):
        """fused=True: no collective afterwards -- the sweep kernel stores every voxel of its slice into all ranks'
        volumes over NVLink (PeerVolume) and a one-element all-reduce on the stream is the only rendezvous."""
        super().__init__()
        self.builder = builder
        self.group = group
        self.fused = fused
        self._peer_volume: Optional[PeerVolume] = None
        self._flag: Optional[Tensor] = None
        # fused path: the sweep kernel's tap tables of this rank's candidate slice are kept across calls
        # (ops.build_tap_table — the first call with a given shape / masks tensor is the BUILD launch)
        self.keep_tap_table = True
        self._tap_cache = None  # (key, uint8 table or None)
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
        if self.fused:
            return self._sweep_fused(feats, g, gm, masks, D)
        if hasattr(self.builder, "sweep_with_valid"):
            local = self.builder.sweep(feats, g, gm, masks)
        else:
            local = self.builder.sweep(feats, g, masks)
        return all_gather_candidates(local, D, self.group)

    def _sweep_fused(self, feats, g, gm, masks, D):
        from . import abi, ops

        rank, world = _rank_world(None, None, self.group)
        d0, d1 = balanced_range(D, rank, world)
        b = self.builder
        std = b._aggregate == abi.AGG_STD
        B, CAM, Cc = feats.shape[:3]
        cout = Cc if std else CAM * Cc
        if self.full_rig is not None:
            Ho, Wo = self.full_rig.cv_camera.H, self.full_rig.cv_camera.W
        else:
            Ho, Wo = g.shape[3:5]
        # channels-last volume (builder.set_vol_channels_last): the peer buffers hold [B, D, Ho, Wo, C] and the module
        # sees the permuted [B, C, D, Ho, Wo] view; a warp's peer store is then 512 contiguous bytes over NVLink
        cl = bool(getattr(b, "vol_channels_last", False))
        shape = (B, D, Ho, Wo, cout) if cl else (B, cout, D, Ho, Wo)
        if self._peer_volume is None or self._peer_volume.shape != shape:
            if self._peer_volume is not None:
                self._peer_volume.close()
            self._peer_volume = PeerVolume(shape, feats.device, self.group)
            self._flag = torch.zeros(1, device=feats.device)
        vol, peers = self._peer_volume.next()
        if cl:
            vol = vol.permute(0, 4, 1, 2, 3)
        mask_hw = None
        if std and b._static_masks is not None:  # the module's packed bits: no packing kernel per call
            masks, mask_hw = b._packed_masks(b._static_masks, check=False)
        call = ops.SweepCall(feats, g, gm, masks if std else None, b._rig_blob, b._aggregate, Ho, Wo, d1 - d0,
                             out=vol, cand_offset=d0, peer_ptrs=peers, channels_last=cl, mask_hw=mask_hw)
        ran = False
        if self.keep_tap_table and g is None and not torch.cuda.is_current_stream_capturing():
            m = masks if std else None
            key = (tuple(feats.shape), tuple(feats.stride()), None if m is None else (m.data_ptr(), m._version, mask_hw),
                   shape, d0, d1, feats.device)
            c = self._tap_cache
            if c is not None and c[0] == key:
                if c[1] is not None:
                    call.set_tap_table(c[1])
            else:
                # BUILD is a complete sweep of this call (peer stores included) that also keeps its records
                table = ops.build_tap_table(call)
                self._tap_cache = (key, table)
                ran = table is not None
        if not ran:
            call.run()
        if world > 1:
            # stream-ordered rendezvous: it starts after this rank's sweep kernel and ends when every rank has joined,
            # i.e. when every rank's stores into this volume have been issued and their kernels have completed
            dist.all_reduce(self._flag, group=self.group)
        return vol

    def forward(self, feats, grids, grid_masks, masks):
        return self.builder.post_vol(self.sweep(feats, grids, grid_masks, masks))


def shard_lists(n: int, world: int) -> List[Tuple[int, int]]:
    """All ranks' ranges, for logging and tests."""
    return [balanced_range(n, r, world) for r in range(world)]
