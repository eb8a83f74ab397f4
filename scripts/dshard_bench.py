#!/usr/bin/env python
"""Candidate-sharded sweep (BASELINE.json config 5, high-D stress) on N GPUs of one box: correctness + timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dshard_bench.py [--workload HIGHD] [--batch 1] [--steps 10]

Every rank holds the full batch of features; rank r sweeps its slice of the candidate list (fused in-kernel projection
from a sub-rig) and one NCCL all-gather + strided copy re-assembles [B, C, D, Ho, Wo] on every rank
(mvs_gi_b200.sharding.CandidateShardedSweep).  Rank 0 also runs the un-sharded sweep and checks bit-equality.
Prints one JSON line: local sweep ms, all-gather + re-assembly ms, un-sharded ms, NVLink GB/s received per rank.
"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="HIGHD")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--small", action="store_true", help="shrunken shapes (tests)")
    ap.add_argument("--channels-last", action="store_true", help="channels_last_3d volume (512-byte peer stores)")
    args = ap.parse_args()
    from mvs_gi_b200 import modules, sharding, synthetic

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = synthetic.WORKLOADS[args.workload]
    if args.small:
        w = synthetic.small(w, feat_hw=(32, 128), mask_hw=(128, 512), cv_hw=(20, 80), n_cands=7)
    rig = w.rig()
    B = args.batch
    gen = torch.Generator(device=dev).manual_seed(4321)  # same features on every rank
    feats = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)[None]).to(dev)

    full = modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval().attach_rig(rig)
    shard = sharding.CandidateShardedSweep(
        modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval(), rig)
    fused = sharding.CandidateShardedSweep(
        modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval(), rig, fused=True)
    if args.channels_last:
        for m in (full, shard.builder, fused.builder):
            m.set_vol_channels_last(True)
    d0, d1 = sharding.candidate_shard(w.num_cands)

    def timed(fn, n):
        for _ in range(3):
            out = fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    with torch.no_grad():
        vol_sharded, ms_sharded = timed(lambda: shard.sweep(feats, None, None, masks), args.steps)
        local_vol, ms_local = timed(lambda: shard.builder.sweep(feats, None, None, masks), args.steps)
        _, ms_gather = timed(lambda: sharding.all_gather_candidates(local_vol, w.num_cands), args.steps)
        vol_full, ms_full = timed(lambda: full.sweep(feats, None, None, masks), args.steps)
        vol_fused, ms_fused = timed(lambda: fused.sweep(feats, None, None, masks), args.steps)
        # local compute of the fused call without its peers' traffic cannot be separated; time the rendezvous alone
        flag = torch.zeros(1, device=dev)
        _, ms_rdv = timed(lambda: dist.all_reduce(flag), args.steps)
    same = bool(torch.equal(vol_sharded, vol_full)) and bool(torch.equal(vol_fused, vol_full))
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        vol_bytes = vol_full.numel() * 4
        recv = vol_bytes * (world - 1) / world
        print(json.dumps({
            "what": "candidate-sharded sweep", "workload": w.name, "channels_last": bool(args.channels_last), "n_gpus": world, "batch": B,
            "bit_identical_to_unsharded_on_all_ranks": bool(ok.item()),
            "ms_unsharded_one_gpu": ms_full, "ms_sharded_total": ms_sharded, "ms_local_slice_sweep": ms_local,
            "ms_allgather_and_reassemble": ms_gather, "ms_fused_p2p_total": ms_fused, "ms_rendezvous_allreduce": ms_rdv,
            "nvlink_send_gbs_per_rank_fused": recv / (ms_fused * 1e-3) / 1e9, "vol_mbytes": vol_bytes / 1e6,
            "nvlink_recv_gbs_per_rank": recv / (ms_gather * 1e-3) / 1e9,
            "candidates_rank0": [d0, d1],
        }), flush=True)
    if fused._peer_volume is not None:
        fused._peer_volume.close()
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
