"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) `gloo` process groups.

The device work is stood in for by the CPU oracle (allowed here: tests/ is one of the places that may call oracle/),
so what is under test is mvs_gi_b200.sharding: the frame partition (no collective), the candidate partition and the
all-gather re-assembly of candidate slices into the [B, C, D, Ho, Wo] layout.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvs_gi_b200 import sharding, synthetic
from oracle import oracle as orc


def test_balanced_range_covers_everything_once():
    for n in (0, 1, 5, 16, 64, 127):
        for world in (1, 2, 3, 4, 8):
            pieces = sharding.shard_lists(n, world)
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sharding.balanced_range(4, 2, 2)


def test_sub_rig_slices_candidates():
    w = synthetic.small(synthetic.WORKLOADS["G16V"])
    rig = w.rig()
    sub = sharding.sub_rig(rig, 4, 9)
    assert sub.num_cands == 5 and list(sub.dist_list) == list(rig.dist_list[4:9])
    g_full, m_full = orc.make_grids(rig.to_struct())
    g_sub, m_sub = orc.make_grids(sub.to_struct())
    assert np.array_equal(g_sub, g_full[:, 4:9]) and np.array_equal(m_sub, m_full[:, 4:9])
    with pytest.raises(ValueError):
        sharding.sub_rig(rig, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_cands, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 4
        w = synthetic.small(synthetic.WORKLOADS["G16V"], feat_hw=(16, 64), mask_hw=(64, 256), cv_hw=(6, 24),
                            n_cands=n_cands)
        rig = w.rig()
        feats = synthetic.make_feats(w, B, 5)
        masks = synthetic.erase_rectangles(synthetic.fisheye_footprint_masks(w), B, 6)
        g, gm = orc.make_grids(rig.to_struct())
        full = orc.sweep_std(feats, g[None], gm[None], masks)

        # ---- frame-sharded: each rank sweeps its own frames, nothing is exchanged
        f0, f1 = sharding.frame_shard(B)
        mine = orc.sweep_std(feats[f0:f1], g[None], gm[None], masks[f0:f1]) if f1 > f0 else full[:0]
        ok_frames = np.array_equal(mine, full[f0:f1])

        # ---- candidate-sharded: sub-rig slice + all-gather re-assembly
        d0, d1 = sharding.candidate_shard(n_cands)
        sub = sharding.sub_rig(rig, d0, d1)
        gs, gms = orc.make_grids(sub.to_struct())
        local = orc.sweep_std(feats, gs[None], gms[None], masks)
        out = sharding.all_gather_candidates(torch.from_numpy(local), n_cands)
        ok_cands = np.array_equal(out.numpy(), full)

        # wrong slice width is rejected before any collective is issued
        try:
            sharding.all_gather_candidates(torch.zeros(1, 1, (d1 - d0) + 1, 2, 2), n_cands)
            ok_err = False
        except RuntimeError:
            ok_err = True
        ret[rank] = (ok_frames, ok_cands, ok_err, (f0, f1), (d0, d1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cands", [(2, 8), (2, 7), (3, 8)])
def test_sharding_over_gloo(world, n_cands):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_cands, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
        for p in procs:
            if p.is_alive():
                p.kill()
                pytest.fail("gloo worker hung")
            assert p.exitcode == 0
        got = dict(ret)
    assert sorted(got) == list(range(world))
    for r in range(world):
        ok_frames, ok_cands, ok_err, fr, dr = got[r]
        assert ok_frames, f"rank {r}: frame shard {fr} differs from the un-sharded sweep"
        assert ok_cands, f"rank {r}: all-gathered candidate slices differ from the un-sharded sweep"
        assert ok_err
    assert [got[r][3] for r in range(world)] == sharding.shard_lists(4, world)
    assert [got[r][4] for r in range(world)] == sharding.shard_lists(n_cands, world)
