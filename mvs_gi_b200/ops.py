"""PyTorch custom ops over the C ABI (include/mvsgi_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every op below
marshals raw pointers + strides into the POD argument structs and enqueues the
hand-written sm_100a kernels through ctypes.  There is no fallback path: a
non-CUDA tensor or a missing library is an error.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import abi


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"mvs_gi_b200: `{name}` must be a CUDA tensor (got {t.device}); there is no CPU path")


def rig_to_blob(rig_struct: abi.MvsgiRig) -> Tensor:
    """MvsgiRig -> CPU uint8 tensor (how calibration travels through torch.library)."""
    raw = bytes(rig_struct)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()


_RIG_CACHE: dict = {}  # (data_ptr, version, numel) of a rig blob -> (weak reference to the blob, decoded struct)


def _blob_to_rig(blob: Tensor) -> abi.MvsgiRig:
    """Decode a calibration blob; the struct of a blob already seen (same tensor, same version — a module's attached rig)
    is reused, so the per-call host work of the single-frame path does not include re-parsing ~800 bytes."""
    import weakref

    key = (blob.data_ptr(), blob._version, blob.numel())
    hit = _RIG_CACHE.get(key)
    if hit is not None and hit[0]() is blob:
        return hit[1]
    buf = blob.contiguous().numpy().tobytes()
    if len(buf) != C.sizeof(abi.MvsgiRig):
        raise RuntimeError("rig blob has the wrong size for this ABI")
    rig = abi.MvsgiRig.from_buffer_copy(buf)
    if len(_RIG_CACHE) > 64:
        _RIG_CACHE.clear()
    try:
        _RIG_CACHE[key] = (weakref.ref(blob), rig)
    except TypeError:
        pass
    return rig


def _fill_strides(dst, strides) -> None:
    for i, s in enumerate(strides):
        dst[i] = int(s)


class SweepCall:
    """One prepared sweep invocation: argument struct, outputs and workspace allocated once, then `run()` enqueues.

    The custom op below builds one per call; bench.py keeps one per input set and calls `run(FLAG_STAGE_ONLY)` /
    `run(FLAG_SWEEP_ONLY)` so that CUDA events can bracket the dominant kernel on its own."""

    def __init__(self, feats: Tensor, grids: Optional[Tensor], grid_masks: Optional[Tensor], masks: Optional[Tensor],
                 rig_blob: Optional[Tensor], aggregate: int, out_h: int, out_w: int, num_cands: int,
                 want_valid: bool = False, out: Optional[Tensor] = None, cand_offset: int = 0,
                 peer_ptrs: Sequence[int] = (), describe_only: bool = False, channels_last: bool = False,
                 feats_bf16: bool = False, cta_table: bool = False, staged_feats: bool = False,
                 mask_hw: Optional[Tuple[int, int]] = None, vol_bf16: bool = False,
                 octet_order: Optional[Tensor] = None):
        """`out`, `cand_offset`, `peer_ptrs`: candidate-sharded sweep.  `out` is the full [B, Cout, Dvol, Ho, Wo]
        volume (contiguous); this call fills candidates [cand_offset, cand_offset + num_cands) of it and of every
        peer-mapped volume in `peer_ptrs` (raw device pointers from sharding.PeerVolume).
        `channels_last`: the volume is a [B, Cout, D, Ho, Wo] tensor with torch.channels_last_3d strides (memory
        [B, D, Ho, Wo, Cout], MVSGI_FLAG_VOL_CHANNELS_LAST) — what post_vol's Conv3d takes without a transpose.
        `feats_bf16`: stage the features as bfloat16 (MVSGI_FLAG_FEATS_BF16, stated-tolerance mode, fp32 arithmetic).
        `cta_table`: MVSGI_FLAG_CTA_TABLE (only meaningful in experiment builds of the library, which also hold the
        tile-walking kernel).
        `masks` may be the int32 tensor `pack_masks()` returns ([B|1, CAM, words], MVSGI_MASK_PACKED) together with
        `mask_hw` = (Hm, Wm) of the masks it was packed from: no packing kernel, no float masks read.
        `staged_feats`: stage channels-last features into the zero-framed copy instead of tapping them in place
        (MVSGI_FLAG_STAGED_FEATS; same volume, for tests and bench.py).
        `vol_bf16`: the volume is written as bfloat16 (MVSGI_FLAG_VOL_BF16, needs `channels_last`; stated-tolerance
        mode for a bf16 post_vol on the tensor cores).
        `octet_order`: the int32 [runs, units] table of `balanced_octet_order()` (cost-balanced assignment of the
        image's 8-pixel octets to the resident CTAs; same volume bit for bit).
        `describe_only`: fill the argument struct (inputs only) without allocating the volume or the workspace —
        how the backward op describes the forward problem to mvsgi_sweep_backward_f32."""
        self.lib = abi.load_library()
        _require_cuda(feats, "feats")
        if feats.dtype != torch.float32 or feats.dim() != 5:
            raise RuntimeError("feats must be float32 [B, CAM, C, H, W]")
        B, CAM, Cc, Hf, Wf = feats.shape
        a = abi.MvsgiSweepArgs()
        a.struct_size = C.sizeof(abi.MvsgiSweepArgs)
        a.aggregate = aggregate
        self.base_flags = (abi.FLAG_VOL_CHANNELS_LAST if channels_last else 0) | \
                          (abi.FLAG_FEATS_BF16 if feats_bf16 else 0) | (abi.FLAG_CTA_TABLE if cta_table else 0) | \
                          (abi.FLAG_STAGED_FEATS if staged_feats else 0) | (abi.FLAG_VOL_BF16 if vol_bf16 else 0)
        if vol_bf16 and not channels_last:
            raise RuntimeError("a bfloat16 volume needs the channels-last layout (channels_last=True)")
        a.flags = self.base_flags
        a.batch, a.num_cams, a.feat_chs, a.num_cands = B, CAM, Cc, num_cands
        a.feat_h, a.feat_w, a.out_h, a.out_w = Hf, Wf, out_h, out_w
        a.feats = feats.data_ptr()
        _fill_strides(a.feats_stride, feats.stride())
        self.keep = [feats]
        self.device = feats.device

        if aggregate == abi.AGG_STD:
            if masks is None:
                raise RuntimeError("the std builder needs masks")
            _require_cuda(masks, "masks")
            if masks.dtype == torch.int32:  # bits from pack_masks()
                if mask_hw is None or masks.dim() != 3 or masks.shape[1] != CAM or masks.shape[0] not in (1, B):
                    raise RuntimeError("packed masks must be the int32 [B|1, CAM, words] tensor of pack_masks() plus mask_hw")
                if masks.stride(2) != 1:
                    raise RuntimeError("packed masks must be contiguous in their last dimension")
                a.masks_dtype = abi.MASK_PACKED
                a.mask_h, a.mask_w = int(mask_hw[0]), int(mask_hw[1])
                a.masks = masks.data_ptr()
                _fill_strides(a.masks_stride, (0 if masks.shape[0] == 1 else masks.stride(0), masks.stride(1), 0, 0))
            else:
                if masks.dim() != 5 or masks.shape[1] != CAM or masks.shape[2] != 1 or masks.shape[0] not in (1, B):
                    raise RuntimeError(f"masks must be [B|1, CAM, 1, Hm, Wm], got {tuple(masks.shape)}")
                if masks.dtype == torch.float32:
                    a.masks_dtype = abi.MASK_F32
                elif masks.dtype in (torch.bool, torch.uint8):
                    a.masks_dtype = abi.MASK_U8
                else:
                    raise RuntimeError(f"masks dtype {masks.dtype} not supported (float32, bool, uint8)")
                a.mask_h, a.mask_w = masks.shape[3], masks.shape[4]
                a.masks = masks.data_ptr()
                ms = masks.stride()
                _fill_strides(a.masks_stride, (0 if masks.shape[0] == 1 else ms[0], ms[1], ms[3], ms[4]))
            self.keep.append(masks)

        self.rig = None
        if rig_blob is not None:
            a.grid_mode = abi.GRID_PROJECT
            self.rig = _blob_to_rig(rig_blob)
            a.rig = C.pointer(self.rig)
        else:
            a.grid_mode = abi.GRID_SUPPLIED
            if grids is None:
                raise RuntimeError("grids are required when no rig calibration is given")
            _require_cuda(grids, "grids")
            if grids.dtype != torch.float32 or grids.dim() != 6 or grids.shape[-1] != 2 or grids.shape[1] != CAM:
                raise RuntimeError(f"grids must be float32 [B|1, CAM, D, Ho, Wo, 2], got {tuple(grids.shape)}")
            if grids.shape[0] not in (1, B):
                raise RuntimeError("grids batch dimension must be 1 or B")
            if grids.stride(-1) != 1:
                grids = grids.contiguous()
            gs = grids.stride()
            a.grids = grids.data_ptr()
            _fill_strides(a.grids_stride, (0 if grids.shape[0] == 1 else gs[0], gs[1], gs[2], gs[3], gs[4]))
            self.keep.append(grids)
            if aggregate == abi.AGG_STD:
                if grid_masks is None:
                    raise RuntimeError("the std builder needs grid_masks")
                _require_cuda(grid_masks, "grid_masks")
                gm = grid_masks
                if gm.dim() == 6:
                    gm = gm[..., 0]
                if gm.dtype == torch.bool:
                    gm = gm.view(torch.uint8)
                elif gm.dtype != torch.uint8:
                    gm = (gm != 0).view(torch.uint8)
                if tuple(gm.shape[1:]) != tuple(grids.shape[1:5]) or gm.shape[0] not in (1, B):
                    raise RuntimeError("grid_masks shape does not match grids")
                a.grid_masks = gm.data_ptr()
                gms = gm.stride()
                _fill_strides(a.grid_masks_stride,
                              (0 if gm.shape[0] == 1 else gms[0], gms[1], gms[2], gms[3], gms[4]))
                self.keep.append(gm)

        cout = Cc if aggregate == abi.AGG_STD else CAM * Cc
        self.args = a
        if describe_only:
            return
        if out is None:
            if cand_offset or len(peer_ptrs):
                raise RuntimeError("a candidate-sharded sweep needs the full volume passed as `out`")
            self.vol = torch.empty((B, cout, num_cands, out_h, out_w),
                                   dtype=torch.bfloat16 if vol_bf16 else torch.float32, device=feats.device,
                                   memory_format=torch.channels_last_3d if channels_last else torch.contiguous_format)
        else:
            _require_cuda(out, "out")
            fmt = torch.channels_last_3d if channels_last else torch.contiguous_format
            if vol_bf16:
                raise RuntimeError("a candidate-sharded sweep writes float32 volumes")
            if (out.dtype != torch.float32 or out.dim() != 5 or not out.is_contiguous(memory_format=fmt)
                    or out.shape[0] != B or out.shape[1] != cout or tuple(out.shape[3:]) != (out_h, out_w)):
                raise RuntimeError(f"`out` must be a float32 [B, {cout}, Dvol, {out_h}, {out_w}] volume, "
                                   f"{'channels_last_3d' if channels_last else 'contiguous'}")
            if cand_offset < 0 or cand_offset + num_cands > out.shape[2]:
                raise RuntimeError("candidate slice does not fit the output volume")
            if len(peer_ptrs) > abi.MAX_PEERS:
                raise RuntimeError(f"at most {abi.MAX_PEERS} peer volumes")
            self.vol = out
            a.cand_offset, a.vol_num_cands, a.num_peers = cand_offset, out.shape[2], len(peer_ptrs)
            for i, ptr in enumerate(peer_ptrs):
                a.peer_vols[i] = int(ptr)
        a.vol = self.vol.data_ptr()
        if want_valid and aggregate == abi.AGG_STD:
            if out is not None:
                raise RuntimeError("the validity output is not available for a candidate-sharded sweep")
            self.valid = torch.empty((B, num_cands, out_h, out_w), dtype=torch.uint8, device=feats.device)
            a.valid_out = self.valid.data_ptr()
        else:
            self.valid = torch.empty((0,), dtype=torch.uint8, device=feats.device)
        need = self.lib.mvsgi_sweep_workspace_bytes(C.byref(a))
        if need < 0:
            abi.check(int(need), "mvsgi_sweep_workspace_bytes")
        self.workspace = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=feats.device)
        a.workspace = self.workspace.data_ptr()
        a.workspace_bytes = self.workspace.numel()
        self.args = a
        if octet_order is not None:
            self.set_octet_order(octet_order)

    def set_octet_order(self, order: Optional[Tensor]) -> None:
        """Attach (or, with None, detach) the table of `balanced_octet_order()`."""
        a = self.args
        if order is None:
            a.octet_order, a.order_runs, a.order_units = None, 0, 0
            return
        _require_cuda(order, "octet_order")
        if order.dtype != torch.int32 or order.dim() != 2 or not order.is_contiguous():
            raise RuntimeError("octet_order must be a contiguous int32 [runs, units] tensor")
        a.octet_order, a.order_runs, a.order_units = order.data_ptr(), order.shape[0], order.shape[1]
        self.keep.append(order)

    def tap_table_bytes(self) -> int:
        """Bytes of the tap table this problem keeps across calls (mvsgi_sweep_tap_table_bytes); 0 = none (per-frame
        grids or per-frame masks)."""
        n = int(self.lib.mvsgi_sweep_tap_table_bytes(C.byref(self.args)))
        if n < 0:
            abi.check(n, "mvsgi_sweep_tap_table_bytes")
        return n

    def set_tap_table(self, table: Optional[Tensor], mode: int = abi.TAP_TABLE_USE) -> None:
        """Attach (or, with None, detach) the uint8 buffer of `build_tap_table()` (MvsgiSweepArgs.tap_table)."""
        a = self.args
        if table is None:
            a.tap_table, a.tap_table_bytes, a.tap_table_mode = None, 0, abi.TAP_TABLE_OFF
            return
        _require_cuda(table, "tap_table")
        if table.dtype != torch.uint8 or table.dim() != 1 or not table.is_contiguous():
            raise RuntimeError("tap_table must be the contiguous uint8 buffer of build_tap_table()")
        a.tap_table, a.tap_table_bytes, a.tap_table_mode = table.data_ptr(), table.numel(), mode
        self.keep.append(table)

    def launch_count(self, flags: int = 0) -> int:
        self.args.flags = flags | self.base_flags
        n = self.lib.mvsgi_sweep_launch_count(C.byref(self.args))
        self.args.flags = self.base_flags
        return int(n)

    def run(self, flags: int = 0) -> None:
        self.args.flags = flags | self.base_flags
        with torch.cuda.device(self.device):
            rc = self.lib.mvsgi_sweep_f32(C.byref(self.args), _stream_ptr(self.device))
        self.args.flags = self.base_flags
        abi.check(rc, "mvsgi_sweep_f32")


def lpt_assign(cost, runs: int, units: int, slowdown=None):
    """Deal octets (index = position in `cost`) to `runs` runs of at most `units` slots so that the runs' total costs
    are as equal as their equal cardinality allows: longest-processing-time-first in rounds — the octets in descending
    cost, `runs` at a time, the dearest of a round to the run that is lightest so far — so every run gets one octet per
    round and the spread left after a round is at most the spread of costs inside it.
    `slowdown` [runs]: relative time per unit of cost of the CTA that will sweep each run (1 = nominal; a run on an SM
    that takes 8 % longer has 1.08): the loads are equalised in time, load * slowdown, instead of in cost.
    -> int32 [runs, units], -1 = empty slot; every octet appears exactly once."""
    import numpy as np

    cost = np.asarray(cost, dtype=np.int64)
    n = int(cost.size)
    if runs * units < n:
        raise ValueError("more octets than slots")
    slow = np.ones(runs) if slowdown is None else np.asarray(slowdown, dtype=np.float64)
    if slow.shape != (runs,) or not np.all(slow > 0):
        raise ValueError("slowdown must hold one positive factor per run")
    order = np.full((runs, units), -1, dtype=np.int32)
    load = np.zeros(runs, dtype=np.float64)
    desc = np.argsort(-cost, kind="stable")
    for k in range(units):
        chunk = desc[k * runs:(k + 1) * runs]
        if chunk.size == 0:
            break
        # lightest runs first; a slow run counts as if it already carried its handicap over the rounds still to come
        ahead = load + (slow - 1.0) * (units - k) * float(cost[chunk].mean()) if slowdown is not None else load
        light = np.argsort(ahead, kind="stable")[:chunk.size]
        order[light, k] = chunk
        load[light] += cost[chunk] * slow[light]
    # compact: empty slots (only possible in the last, partial round) stay at the end of their run by construction
    return order


def build_tap_table(call: "SweepCall") -> Optional[Tensor]:
    """The sweep kernel's tap tables for the problem `call` describes, kept for later calls (MvsgiSweepArgs.tap_table):
    one sweep of `call` in BUILD mode (stages / packs what the call needs, builds the records as usual and stores them),
    after which `call` — and every call with the same shapes, strides, masks, rig and octet_order — copies them in with
    one bulk asynchronous copy per work item instead of evaluating the geometry.  The role of the grids the reference
    precomputes once per rig (multi_view_camera_model_dataset.py:519-557).  Attach `octet_order` first: the table
    follows the work assignment.  -> uint8 buffer (attached to `call` in USE mode), or None when the problem keeps no
    table (per-frame grids / masks)."""
    n = call.tap_table_bytes()
    if n <= 0:
        return None
    # zero-filled: slots of items with fewer records than the largest are copied whole-sale up to their own length only,
    # but a defined buffer keeps compute-sanitizer's initcheck quiet
    table = torch.zeros((n,), dtype=torch.uint8, device=call.device)
    call.set_tap_table(table, abi.TAP_TABLE_BUILD)
    call.run()
    call.set_tap_table(table, abi.TAP_TABLE_USE)
    return table


def balanced_octet_order(call: "SweepCall") -> Optional[Tensor]:
    """Cost-balanced assignment of the image's octets to the sweep's resident CTAs for the problem `call` describes
    (static calibration, packed static masks): the per-octet cost comes from the device (mvsgi_sweep_octet_cost, the
    cameras each voxel will sample), the assignment is longest-processing-time-first on the host.  One device
    synchronisation — meant to be done once per (rig, masks, shape) and cached, like the packed masks.  None when the
    problem is not swept as one item per resident CTA (high-D, per-frame grids): the default assignment stays."""
    lib = call.lib
    runs, units, noct, ctas = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.mvsgi_sweep_balance_shape(C.byref(call.args), C.byref(runs), C.byref(units), C.byref(noct), C.byref(ctas))
    if rc == abi.E_UNSUPPORTED:
        return None
    abi.check(rc, "mvsgi_sweep_balance_shape")
    cost = torch.empty((noct.value,), dtype=torch.int32, device=call.device)
    with torch.cuda.device(call.device):
        rc = lib.mvsgi_sweep_octet_cost(C.byref(call.args), cost.data_ptr(), _stream_ptr(call.device))
    if rc == abi.E_UNSUPPORTED:
        return None
    abi.check(rc, "mvsgi_sweep_octet_cost")
    order = deal_octets(cost.cpu().numpy(), runs.value, units.value, ctas.value)
    return torch.from_numpy(order).to(call.device)


def deal_octets(cost, runs: int, units: int, ctas: int, slowdown=None):
    """The [runs, units] table for `ctas` CTAs that each sweep runs c, c + ctas, ...: lpt_assign over the CTAs (a CTA's
    time is the sum over its runs; `slowdown` per CTA), each CTA's octets then cut into its runs in order."""
    import numpy as np

    if runs <= ctas:
        return lpt_assign(cost, runs, units, slowdown)
    if runs % ctas != 0:
        raise ValueError("runs must be whole waves of the CTAs")
    k = runs // ctas
    per_cta = lpt_assign(cost, ctas, k * units, slowdown)          # [ctas, k * units], -1 only at the end of a row
    return np.ascontiguousarray(per_cta.reshape(ctas, k, units).transpose(1, 0, 2).reshape(runs, units))


def cta_times(call: "SweepCall", launches: int = 3):
    """(sm [ctas], microseconds [launches, ctas]) of the sweep kernel's CTAs for `call` (MvsgiSweepArgs.cta_times_out),
    after one warm-up launch.  Synchronises."""
    import numpy as np

    n = call.lib.mvsgi_sweep_grid(C.byref(call.args))
    if n < 0:
        abi.check(int(n), "mvsgi_sweep_grid")
    buf = torch.zeros((n, 4), dtype=torch.int64, device=call.device)
    call.args.cta_times_out = buf.data_ptr()
    try:
        call.run()
        out = []
        for _ in range(launches):
            call.run()
            torch.cuda.synchronize(call.device)
            t = buf.cpu().numpy()
            out.append((t[:, 3] - t[:, 1]) / 1e3)
        sm = t[:, 0].copy()
    finally:
        call.args.cta_times_out = None
    return sm, np.stack(out)


def calibrated_octet_order(call: "SweepCall", rounds: int = 3, threshold: float = 0.015, launches: int = 3):
    """balanced_octet_order(), then a few measured corrections: the sweep is run with the table and per-CTA timestamps
    (MvsgiSweepArgs.cta_times_out); a CTA that finishes later than the median — on the B200s measured the same handful
    of CTA slots does, launch after launch, by 5-8 %: the placement of a full one-wave grid on the SMs repeats — gets
    the ratio as the `slowdown` of its run, and the octets are dealt again (ops.lpt_assign).  Up to `rounds` such
    steps; the table with the shortest measured kernel time is returned, with a small report — never worse than the
    balanced table it starts from.  `call` must be a complete SweepCall (its volume is overwritten); a dozen launches
    and synchronisations — once per (device, rig, masks, shape), outside any timed or captured region."""
    import numpy as np

    order = balanced_octet_order(call)
    if order is None:
        return None, {"calibrated": False, "why": "the problem is not swept as one item per resident CTA"}
    saved = (call.args.octet_order, call.args.order_runs, call.args.order_units)
    try:
        runs, units = order.shape
        noct = int((order >= 0).sum())
        cost_t = torch.empty((noct,), dtype=torch.int32, device=call.device)
        with torch.cuda.device(call.device):
            abi.check(call.lib.mvsgi_sweep_octet_cost(C.byref(call.args), cost_t.data_ptr(), _stream_ptr(call.device)),
                      "mvsgi_sweep_octet_cost")
        cost = cost_t.cpu().numpy()
        call.set_octet_order(order)
        _, us = cta_times(call, launches)
        ctas = us.shape[1]
        best, best_us = order, float(us.max(axis=1).min())
        report = {"calibrated": False, "kernel_us_balanced": best_us, "kernel_us_rounds": []}
        slow = np.ones(ctas)
        for _ in range(rounds):
            med = np.median(us, axis=0)                   # per CTA, over the launches
            ratio = med / float(np.median(med))
            step = np.where(np.abs(ratio - 1.0) > threshold, np.clip(ratio, 0.95, 1.12), 1.0)
            if np.all(step == 1.0):
                break
            slow = slow * step
            slow = slow / np.median(slow)
            cand = torch.from_numpy(deal_octets(cost, runs, units, ctas, slow)).to(call.device)
            call.set_octet_order(cand)
            _, us = cta_times(call, launches)
            k_us = float(us.max(axis=1).min())
            report["kernel_us_rounds"].append(k_us)
            if k_us < best_us:
                best, best_us = cand, k_us
                report["calibrated"] = True
        report["kernel_us"] = best_us
        report["ctas_slower_than_2pct"] = int((slow > 1.02).sum())
        return best, report
    finally:
        call.args.octet_order, call.args.order_runs, call.args.order_units = saved


def _sweep_impl(feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands, want_valid,
                channels_last=False, feats_bf16=False, mask_h=0, mask_w=0, vol_bf16=False, octet_order=None,
                tap_table=None):
    call = SweepCall(feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands, want_valid,
                     channels_last=channels_last, feats_bf16=feats_bf16,
                     mask_hw=(mask_h, mask_w) if mask_h > 0 else None, vol_bf16=vol_bf16, octet_order=octet_order)
    if tap_table is not None:
        call.set_tap_table(tap_table, abi.TAP_TABLE_USE)
    call.run()
    # inputs and workspace are only read by kernels already enqueued on this stream; the caching allocator reuses
    # freed blocks in stream order, so dropping the references here is safe.
    return call.vol, call.valid


@torch.library.custom_op("mvsgi::sweep", mutates_args=())
def sweep_op(feats: Tensor, grids: Optional[Tensor], grid_masks: Optional[Tensor], masks: Optional[Tensor],
             rig_blob: Optional[Tensor], aggregate: int, out_h: int, out_w: int, num_cands: int,
             want_valid: bool, channels_last: bool = False, feats_bf16: bool = False, mask_h: int = 0,
             mask_w: int = 0, vol_bf16: bool = False, octet_order: Optional[Tensor] = None,
             tap_table: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """`masks`: [B|1, CAM, 1, Hm, Wm] float32 / bool / uint8 (reference interface), or the int32 bits of pack_masks()
    with mask_h, mask_w = the size of the masks they were packed from.  `octet_order`: balanced_octet_order().
    `tap_table`: build_tap_table() of a call with exactly these shapes, strides, masks, rig and octet_order."""
    return _sweep_impl(feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands, want_valid,
                       channels_last, feats_bf16, mask_h, mask_w, vol_bf16, octet_order, tap_table)


@sweep_op.register_fake
def _(feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands, want_valid, channels_last=False,
      feats_bf16=False, mask_h=0, mask_w=0, vol_bf16=False, octet_order=None, tap_table=None):
    B, CAM, Cc = feats.shape[:3]
    cout = Cc if aggregate == abi.AGG_STD else CAM * Cc
    vol = torch.empty((B, cout, num_cands, out_h, out_w), dtype=torch.bfloat16 if vol_bf16 else feats.dtype,
                      device=feats.device,
                      memory_format=torch.channels_last_3d if channels_last else torch.contiguous_format)
    valid = feats.new_empty((B, num_cands, out_h, out_w) if (want_valid and aggregate == abi.AGG_STD) else (0,),
                            dtype=torch.uint8)
    return vol, valid


@torch.library.custom_op("mvsgi::softargmin", mutates_args=())
def softargmin_op(costs: Tensor, inv_dist_idx: Tensor, out_h: int, out_w: int, src_scale: float,
                  want_norm_costs: bool, want_argmax: bool) -> Tuple[Tensor, Tensor, Tensor]:
    lib = abi.load_library()
    _require_cuda(costs, "costs")
    if costs.dtype != torch.float32 or costs.dim() != 5:
        raise RuntimeError("costs must be float32 [B, 1, D, H, W]")
    B, _, D, H, W = costs.shape
    idx = inv_dist_idx.reshape(-1).to(device=costs.device, dtype=torch.float32).contiguous()
    if idx.numel() != D:
        raise RuntimeError(f"inv_dist_idx has {idx.numel()} entries, costs has D={D}")
    inv = torch.empty((B, 1, out_h, out_w), dtype=torch.float32, device=costs.device)
    nc = torch.empty((B, D, out_h, out_w) if want_norm_costs else (0,), dtype=torch.float32, device=costs.device)
    am = torch.empty((B, out_h, out_w) if want_argmax else (0,), dtype=torch.int32, device=costs.device)
    a = abi.MvsgiSoftArgminArgs()
    a.struct_size = C.sizeof(abi.MvsgiSoftArgminArgs)
    a.batch, a.num_cands, a.in_h, a.in_w, a.out_h, a.out_w = B, D, H, W, out_h, out_w
    a.src_scale_h = a.src_scale_w = src_scale
    a.costs = costs.data_ptr()  # channel 0 of dim 1 (distance_regressor.py:52)
    cs = costs.stride()
    _fill_strides(a.costs_stride, (cs[0], cs[2], cs[3], cs[4]))
    a.inv_dist_idx = idx.data_ptr()
    a.inv_dist = inv.data_ptr()
    a.norm_costs = nc.data_ptr() if want_norm_costs else None
    a.argmax = am.data_ptr() if want_argmax else None
    with torch.cuda.device(costs.device):
        abi.check(lib.mvsgi_softargmin_f32(C.byref(a), _stream_ptr(costs.device)), "mvsgi_softargmin_f32")
    return inv, nc, am


@softargmin_op.register_fake
def _(costs, inv_dist_idx, out_h, out_w, src_scale, want_norm_costs, want_argmax):
    B, _, D = costs.shape[:3]
    return (costs.new_empty((B, 1, out_h, out_w)),
            costs.new_empty((B, D, out_h, out_w) if want_norm_costs else (0,)),
            costs.new_empty((B, out_h, out_w) if want_argmax else (0,), dtype=torch.int32))


# ---- backward ---------------------------------------------------------------------------------
@torch.library.custom_op("mvsgi::sweep_backward", mutates_args=())
def sweep_backward_op(grad_vol: Tensor, feats: Tensor, grids: Optional[Tensor], grid_masks: Optional[Tensor],
                      masks: Optional[Tensor], rig_blob: Optional[Tensor], aggregate: int, out_h: int, out_w: int,
                      num_cands: int, mask_h: int = 0, mask_w: int = 0) -> Tensor:
    """d sum(grad_vol * sweep(feats, ...)) / d feats  (autograd of spherical_sweep_avg.py:38-136 / spherical_sweep.py:38-68)."""
    lib = abi.load_library()
    _require_cuda(grad_vol, "grad_vol")
    fwd = SweepCall(feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands,
                    describe_only=True, mask_hw=(mask_h, mask_w) if mask_h > 0 else None)
    B, CAM, Cc = feats.shape[:3]
    cout = Cc if aggregate == abi.AGG_STD else CAM * Cc
    if tuple(grad_vol.shape) != (B, cout, num_cands, out_h, out_w) or grad_vol.dtype != torch.float32:
        raise RuntimeError(f"grad_vol must be float32 [{B}, {cout}, {num_cands}, {out_h}, {out_w}]")
    grad_vol = grad_vol.contiguous()
    # dense channels-last features (element (b, cam, y, x, c) at unit channel stride): the gradient gets the same layout
    # and the kernel scatters straight into it; anything else is staged and the gradient is NCHW
    B_, CAM_, C_, Hf_, Wf_ = feats.shape
    nhwc = (CAM_ * Hf_ * Wf_ * C_, Hf_ * Wf_ * C_, 1, Wf_ * C_, C_)
    if tuple(feats.stride()) == nhwc and C_ % 16 == 0:
        grad_feats = torch.empty_strided(feats.shape, nhwc, dtype=torch.float32, device=feats.device)
    else:
        grad_feats = torch.empty_like(feats, memory_format=torch.contiguous_format)
    b = abi.MvsgiSweepBwdArgs()
    b.struct_size = C.sizeof(abi.MvsgiSweepBwdArgs)
    b.fwd = C.pointer(fwd.args)
    b.grad_vol = grad_vol.data_ptr()
    b.grad_feats = grad_feats.data_ptr()
    _fill_strides(b.grad_feats_stride, grad_feats.stride())
    need = lib.mvsgi_sweep_backward_workspace_bytes(C.byref(b))
    if need < 0:
        abi.check(int(need), "mvsgi_sweep_backward_workspace_bytes")
    ws = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=feats.device)
    b.workspace, b.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(feats.device):
        abi.check(lib.mvsgi_sweep_backward_f32(C.byref(b), _stream_ptr(feats.device)), "mvsgi_sweep_backward_f32")
    return grad_feats


@sweep_backward_op.register_fake
def _(grad_vol, feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands, mask_h=0, mask_w=0):
    B_, CAM_, C_, Hf_, Wf_ = feats.shape
    nhwc = (CAM_ * Hf_ * Wf_ * C_, Hf_ * Wf_ * C_, 1, Wf_ * C_, C_)
    if tuple(feats.stride()) == nhwc and C_ % 16 == 0:
        return torch.empty_strided(feats.shape, nhwc, dtype=torch.float32, device=feats.device)
    return torch.empty_like(feats, memory_format=torch.contiguous_format)


def _sweep_setup(ctx, inputs, output):
    feats, grids, grid_masks, masks, rig_blob, aggregate, out_h, out_w, num_cands = inputs[:9]
    ctx.save_for_backward(feats, grids, grid_masks, masks, rig_blob)
    mask_h, mask_w = (inputs[12], inputs[13]) if len(inputs) > 13 else (0, 0)
    ctx.meta = (aggregate, out_h, out_w, num_cands, mask_h, mask_w)


def _sweep_bwd(ctx, grad_vol, grad_valid):
    feats, grids, grid_masks, masks, rig_blob = ctx.saved_tensors
    gf = None
    if ctx.needs_input_grad[0] and grad_vol is not None:
        grad_vol = grad_vol.float()  # a bfloat16 volume (vol_bf16) hands back a bfloat16 gradient
        gf = sweep_backward_op(grad_vol, feats, grids, grid_masks, masks, rig_blob, *ctx.meta)
    return (gf,) + (None,) * 16


sweep_op.register_autograd(_sweep_bwd, setup_context=_sweep_setup)


def _softargmin_args(costs: Tensor, idx: Tensor, out_h: int, out_w: int, src_scale: float) -> abi.MvsgiSoftArgminArgs:
    B, _, D, H, W = costs.shape
    a = abi.MvsgiSoftArgminArgs()
    a.struct_size = C.sizeof(abi.MvsgiSoftArgminArgs)
    a.batch, a.num_cands, a.in_h, a.in_w, a.out_h, a.out_w = B, D, H, W, out_h, out_w
    a.src_scale_h = a.src_scale_w = src_scale
    a.costs = costs.data_ptr()
    cs = costs.stride()
    _fill_strides(a.costs_stride, (cs[0], cs[2], cs[3], cs[4]))
    a.inv_dist_idx = idx.data_ptr()
    return a


@torch.library.custom_op("mvsgi::softargmin_backward", mutates_args=())
def softargmin_backward_op(grad_inv: Optional[Tensor], grad_nc: Optional[Tensor], costs: Tensor, inv_dist_idx: Tensor,
                           out_h: int, out_w: int, src_scale: float) -> Tensor:
    """Gradient of softargmin wrt costs (only channel 0 of dim 1 is read by the forward, distance_regressor.py:52)."""
    lib = abi.load_library()
    _require_cuda(costs, "costs")
    B, Cc, D, H, W = costs.shape
    idx = inv_dist_idx.reshape(-1).to(device=costs.device, dtype=torch.float32).contiguous()
    gc = (torch.zeros if Cc > 1 else torch.empty)((B, Cc, D, H, W), dtype=torch.float32, device=costs.device)
    fwd = _softargmin_args(costs, idx, out_h, out_w, src_scale)
    b = abi.MvsgiSoftArgminBwdArgs()
    b.struct_size = C.sizeof(abi.MvsgiSoftArgminBwdArgs)
    b.fwd = C.pointer(fwd)
    keep = []
    for name, g, shape in (("grad_inv_dist", grad_inv, (B, 1, out_h, out_w)), ("grad_norm_costs", grad_nc, (B, D, out_h, out_w))):
        if g is None:
            continue
        if tuple(g.shape) != shape or g.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32 {shape}, got {tuple(g.shape)}")
        g = g.contiguous()
        keep.append(g)
        setattr(b, name, g.data_ptr())
    b.grad_costs = gc.data_ptr()
    gs = gc.stride()
    _fill_strides(b.grad_costs_stride, (gs[0], gs[2], gs[3], gs[4]))
    with torch.cuda.device(costs.device):
        abi.check(lib.mvsgi_softargmin_backward_f32(C.byref(b), _stream_ptr(costs.device)),
                  "mvsgi_softargmin_backward_f32")
    return gc


@softargmin_backward_op.register_fake
def _(grad_inv, grad_nc, costs, inv_dist_idx, out_h, out_w, src_scale):
    return torch.empty_like(costs, memory_format=torch.contiguous_format)


def _softargmin_setup(ctx, inputs, output):
    costs, inv_dist_idx, out_h, out_w, src_scale, want_nc, _ = inputs
    ctx.save_for_backward(costs, inv_dist_idx)
    ctx.meta = (out_h, out_w, src_scale, want_nc)
    try:
        ctx.set_materialize_grads(False)  # an unused output then arrives as None instead of a zero tensor
    except Exception:
        pass


def _softargmin_bwd(ctx, g_inv, g_nc, g_am):
    costs, idx = ctx.saved_tensors
    out_h, out_w, src_scale, want_nc = ctx.meta
    if not ctx.needs_input_grad[0] or (g_inv is None and (g_nc is None or not want_nc)):
        return None, None, None, None, None, None, None
    gc = softargmin_backward_op(g_inv, g_nc if want_nc else None, costs, idx, out_h, out_w, src_scale)
    return gc, None, None, None, None, None, None


softargmin_op.register_autograd(_softargmin_bwd, setup_context=_softargmin_setup)


def project_only(rig_struct: abi.MvsgiRig, device) -> Tuple[Tensor, Tensor]:
    """Dump the grids the fused kernel evaluates: ([CAM,D,Ho,Wo,2] float32, [CAM,D,Ho,Wo,1] bool)."""
    lib = abi.load_library()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("project_only runs on a CUDA device only")
    CAM, D, Ho, Wo = rig_struct.num_cams, rig_struct.num_cands, rig_struct.cv_height, rig_struct.cv_width
    grids = torch.empty((CAM, D, Ho, Wo, 2), dtype=torch.float32, device=device)
    gm = torch.empty((CAM, D, Ho, Wo), dtype=torch.uint8, device=device)
    a = abi.MvsgiProjectArgs()
    a.struct_size = C.sizeof(abi.MvsgiProjectArgs)
    a.rig = C.pointer(rig_struct)
    a.grids = grids.data_ptr()
    a.grid_masks = gm.data_ptr()
    with torch.cuda.device(device):
        abi.check(lib.mvsgi_project_only_f32(C.byref(a), _stream_ptr(device)), "mvsgi_project_only_f32")
    return grids, gm.bool().unsqueeze(-1)


def pack_masks(masks: Tensor, check: bool = False) -> Tensor:
    """[N, CAM, 1, Hm, Wm] float32 / bool / uint8 masks -> the int32 [N, CAM, words] bit images the sweep samples
    (mvsgi_pack_masks; bit set iff mask > 0).  Pass the result as `masks` (with mask_h, mask_w) to sweep_op / SweepCall:
    masks that do not change between calls are then read exactly once.
    `check=True` synchronises and raises if a mask value is negative or NaN — for such masks "some positively weighted
    tap is set" is not the reference's `bilinear(mask) > 0` (spherical_sweep_avg.py:92-96), so they are rejected."""
    lib = abi.load_library()
    _require_cuda(masks, "masks")
    if masks.dim() != 5 or masks.shape[2] != 1:
        raise RuntimeError(f"masks must be [N, CAM, 1, Hm, Wm], got {tuple(masks.shape)}")
    if masks.dtype == torch.float32:
        dt = abi.MASK_F32
    elif masks.dtype in (torch.bool, torch.uint8):
        dt = abi.MASK_U8
    else:
        raise RuntimeError(f"masks dtype {masks.dtype} not supported (float32, bool, uint8)")
    N, CAM, _, Hm, Wm = masks.shape
    words = (Hm + 2) * int(lib.mvsgi_mask_pitch_words(Wm)) * 2
    # + 4 words: the sweep's funnel shift may read up to 16 bytes past the last image
    flat = torch.zeros((N * CAM * words + 4,), dtype=torch.int32, device=masks.device)
    flags = torch.zeros((1,), dtype=torch.int32, device=masks.device)
    ms = masks.stride()
    with torch.cuda.device(masks.device):
        abi.check(lib.mvsgi_pack_masks(masks.data_ptr(), dt, N, ms[0], CAM, ms[1], ms[3], ms[4], Hm, Wm, flat.data_ptr(),
                                       flags.data_ptr(), _stream_ptr(masks.device)), "mvsgi_pack_masks")
    if check and int(flags.item()) & 1:
        raise RuntimeError("mvs_gi_b200: masks hold negative or NaN values; the sweep needs masks >= 0 "
                           "(the reference's masks are 0/1, multi_view_camera_model_dataset.py:424-438)")
    return flat[:N * CAM * words].view(N, CAM, words)


def launch_counter() -> int:
    return int(abi.load_library().mvsgi_launch_counter())
