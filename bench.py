#!/usr/bin/env python
"""bench.py — sweep + cost-volume + soft-argmin frames/s on N B200 (G16V shape) and % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of B synthetic frames per GPU:
    (NCHW -> channels-last staging, mask bit-packing) -> fused sweep/aggregate kernel -> soft-argmin kernel.
Frames are independent, so N GPUs shard frames with no collective (weak scaling: B frames per GPU).
`value`      : whole-job frames/s with inputs resident in HBM (device-timed, max over ranks).
`e2e`        : same metric through the public module API with HOST (pinned) inputs: H2D of feats+costs and D2H of
               inv_dist inside the timed region.
`roofline`   : the dominant kernel (sweep_kernel) timed on its own with CUDA events inside the timed region.
`cpu_baseline`: the PyTorch-eager port of the reference path (oracle/ref_port_torch.py) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="G16V")
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--masks", default="per_frame", choices=["per_frame", "shared"],
                    help="per_frame = reference interface [B,CAM,1,Hm,Wm] float32; shared = static inference masks")
    ap.add_argument("--feats-layout", default="nchw", choices=["nchw", "nhwc"],
                    help="memory layout the feature extractor hands over (reference: NCHW)")
    ap.add_argument("--vol-layout", default="ndhwc", choices=["ncdhw", "ndhwc"],
                    help="memory layout of the cost volume: ncdhw = the reference's; ndhwc = channels_last_3d strides "
                         "(MVSGI_FLAG_VOL_CHANNELS_LAST, what cuDNN's NDHWC Conv3d of post_vol consumes)")
    ap.add_argument("--norm-costs", action="store_true", help="also write norm_costs (training consumers)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the other-layout / bf16-staging kernel timings")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-frames", type=int, default=1, help="frames per step of the CPU reference arm")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic_per_launch():
    """dram bytes (read+write) of one sweep_kernel launch from the committed ncu capture, if there is one."""
    p = ROOT / "profiles" / "sweep_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_throughput(w, frames: int, iters: int, warmup: int):
    """frames/s of the PyTorch-eager port of the reference path on the host cores (all threads)."""
    import numpy as np
    import torch

    from mvs_gi_b200 import synthetic
    from oracle import oracle as orc
    from oracle import ref_port_torch as port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rs = w.rig().to_struct()
    g, gm = orc.make_grids(rs)  # the grids the reference would have precomputed on the CPU
    grids = torch.from_numpy(g)[None].repeat(frames, 1, 1, 1, 1, 1)
    gmask = torch.from_numpy(gm)[None, ..., None].repeat(frames, 1, 1, 1, 1, 1)
    feats = torch.from_numpy(synthetic.make_feats(w, frames, 1234))
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w))[None].repeat(frames, 1, 1, 1, 1)
    costs = torch.from_numpy(synthetic.make_costs(w, frames, 1234))
    idx = torch.tensor([w.bf / d for d in np.array(w.dist_list, dtype=np.float32)], dtype=torch.float32)

    def step():
        with torch.no_grad():
            if w.aggregate == "std":
                port.sweep_std(feats, grids, gmask, masks)
            else:
                port.sweep_concat(feats, grids)
            port.soft_argmin(costs, idx, w.scale, True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    return frames * iters / dt, dt / iters * 1e3, cores


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, ms, cores = cpu_reference_throughput(w, args.ref_frames, args.steps, max(args.warmup, 1))
    sample = f"{args.ref_frames} frame(s) of {w.name} per step, {args.steps} steps, torch-eager port, {cores} threads"
    line = {
        "impl": "reference", "metric": "sweep+cost-volume+softargmin frames/s", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, w):
    return {
        "workload": f"{w.name}: {w.aggregate} cv_builder, {w.num_cams} cams, C={w.feat_chs}, D={w.num_cands}, "
                    f"feats {w.feat_hw[0]}x{w.feat_hw[1]}, masks {w.mask_hw[0]}x{w.mask_hw[1]}, "
                    f"cv {w.cv_hw[0]}x{w.cv_hw[1]}, softargmin x{w.scale}",
        "frames_per_gpu_per_step": args.batch, "global_batch": args.batch * args.gpus,
        "grids": "in-kernel projection (Equirectangular surrogates, fixture rig)",
        "masks": args.masks, "feats_layout": args.feats_layout, "vol_layout": args.vol_layout, "norm_costs": bool(args.norm_costs),
        "parallelism": f"frame-sharded x{args.gpus}, no collective",
        "cache": "3 rotating input sets, each > 126 MB L2",
    }


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, w):
    import torch
    import torch.distributed as dist

    from mvs_gi_b200 import abi, modules, ops, synthetic

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi.load_library()

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    rig = w.rig()
    rig_blob = ops.rig_to_blob(rig.to_struct())
    agg = abi.AGG_STD if w.aggregate == "std" else abi.AGG_CONCAT
    nsets = 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    base_masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)
    if w.aggregate == "std":
        if args.masks == "per_frame":
            masks = torch.from_numpy(synthetic.erase_rectangles(base_masks.cpu().numpy(), B, 77 + rank)).to(dev)
        else:
            masks = base_masks[None]
    else:
        masks = None
    calls, costs_sets = [], []
    for s in range(nsets):
        f = torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
        if args.feats_layout == "nhwc":
            f = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        calls.append(ops.SweepCall(f, None, None, masks, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands,
                                   channels_last=args.vol_layout == "ndhwc"))
        costs_sets.append(torch.randn((B, 1, w.num_cands, *w.cv_hw), device=dev, generator=gen))
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
    Ho, Wo, scale = reg._out_size(*w.cv_hw)

    def step(i, ev=None):
        c = calls[i % nsets]
        c.run(abi.FLAG_STAGE_ONLY)
        if ev is not None:
            ev[0].record()
        c.run(abi.FLAG_SWEEP_ONLY)
        if ev is not None:
            ev[1].record()
        return ops.softargmin_op(costs_sets[i % nsets], reg.inv_dist_idx, Ho, Wo, float(scale), args.norm_costs, False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ops.launch_counter()
    barrier()
    e0.record()
    for i in range(K):
        step(i, evs[i])
    e1.record()
    barrier()
    launches = ops.launch_counter() - launches0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    sweep_ms = sum(a.elapsed_time(b) for a, b in evs) / K
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    value = world * B * K / (ms_total_max * 1e-3)

    # ---- the sweep kernel alone in the other volume layout, and with bf16-staged features (stated-tolerance mode),
    # on the same inputs; each timed launch follows a staging pass over another 400 MB input set (L2 evicted)
    def time_variant(channels_last, feats_bf16):
        oc = ops.SweepCall(calls[0].keep[0], None, None, masks, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands,
                           channels_last=channels_last, feats_bf16=feats_bf16)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(6)]
        for e in ev:
            calls[1].run(abi.FLAG_STAGE_ONLY)
            e[0].record()
            oc.run(abi.FLAG_STAGE_ONLY)
            e[1].record()
            oc.run(abi.FLAG_SWEEP_ONLY)
            e[2].record()
        torch.cuda.synchronize()
        med = lambda v: sorted(v)[len(v) // 2]
        return med([e[0].elapsed_time(e[1]) for e in ev[1:]]), med([e[1].elapsed_time(e[2]) for e in ev[1:]])

    variants = None
    if rank == 0 and not args.no_variants:
        cl = args.vol_layout == "ndhwc"
        _, other_ms = time_variant(not cl, False)
        bf_stage_ms, bf_ms = time_variant(cl, True)
        variants = {"other_vol_layout": {"layout": "ncdhw" if cl else "ndhwc", "kernel_ms": other_ms},
                    "feats_bf16": {"layout": args.vol_layout, "kernel_ms": bf_ms, "stage_ms": bf_stage_ms,
                                   "tolerance": "max|vol - fp32 vol| <= 1e-2 * max|vol| (tests/test_gpu_parity.py)"}}
    barrier()

    # ---- single-frame latency (the reference deploys at batch 1, api/inference_pytorch.py): eager launches and the same
    # four kernels captured once in a CUDA graph (the library neither allocates nor synchronises)
    latency = None
    if rank == 0 and not args.no_variants:
        f1 = calls[0].keep[0][:1].contiguous()
        m1 = masks[:1] if masks is not None else None
        c1 = ops.SweepCall(f1, None, None, m1, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands,
                           channels_last=args.vol_layout == "ndhwc")
        k1 = costs_sets[0][:1].contiguous()

        def one():
            c1.run()
            return ops.softargmin_op(k1, reg.inv_dist_idx, Ho, Wo, float(scale), args.norm_costs, False)

        def time_it(fn, n=200):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        eager_ms = time_it(one)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one()
        graph_ms = time_it(graph.replay)
        latency = {"frames": 1, "ms_eager": eager_ms, "ms_cuda_graph": graph_ms, "launches": 4}
        del graph
    barrier()

    # ---- end to end through the public module API, host buffers
    e2e = None
    if not args.no_e2e:
        builder_cls = modules.SphericalSweepStdMasked if w.aggregate == "std" else modules.SphericalSweep
        builder = builder_cls(w.num_cams, w.out_chs, 3).to(dev).eval().attach_rig(rig)
        if w.aggregate == "std":
            builder.set_static_masks(base_masks[None])  # inference: masks uploaded once (api/inference_class.py:40-45)
        builder.set_vol_channels_last(args.vol_layout == "ndhwc")
        reg.return_norm_costs = bool(args.norm_costs)
        h_feats = [torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw)).pin_memory() for _ in range(2)]
        h_costs = [torch.randn((B, 1, w.num_cands, *w.cv_hw)).pin_memory() for _ in range(2)]
        h_out = torch.empty((B, 1, Ho, Wo)).pin_memory()
        checksum = 0.0

        # double-buffered upload: the H2D copy of step i+1 runs on a copy stream while step i computes; every step's
        # inputs still cross PCIe from pinned host memory inside the timed region and its result goes back to the host
        copy_stream = torch.cuda.Stream()
        d_feats = [torch.empty_like(h_feats[0], device=dev) for _ in range(2)]
        d_costs = [torch.empty_like(h_costs[0], device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            s_ = i % 2
            copy_stream.wait_event(done[s_])  # the step that last read this device buffer has finished
            with torch.cuda.stream(copy_stream):
                d_feats[s_].copy_(h_feats[s_], non_blocking=True)
                d_costs[s_].copy_(h_costs[s_], non_blocking=True)
                ready[s_].record(copy_stream)

        def e2e_run(n):
            cur = torch.cuda.current_stream()
            upload(0)
            vol = None
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                s_ = i % 2
                cur.wait_event(ready[s_])
                with torch.no_grad():
                    f, c = d_feats[s_], d_costs[s_]
                    vol = builder.sweep(f, None, None, None) if w.aggregate == "std" else builder.sweep(f, None)
                    inv, _ = reg(c)
                    h_out.copy_(inv, non_blocking=True)
                done[s_].record(cur)
            return vol

        e2e_run(3)
        barrier()
        Ke = K
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_run(Ke)
        s1.record()
        barrier()
        checksum = float(h_out.sum())
        t2 = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * Ke / (float(t2.item()) * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": int(h_feats[0].numel() * 4 + h_costs[0].numel() * 4),
               "d2h_bytes_per_step": int(h_out.numel() * 4), "steps": Ke, "result_checksum": checksum}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sweep_kernel), algorithmic bytes per launch (SURVEY.md §8d / DESIGN.md)
    peak, peak_kind = measured_peak_gbs()
    mask_bits_bytes = 0
    if w.aggregate == "std":
        n_img = (B if args.masks == "per_frame" else 1) * w.num_cams
        pitch = (w.mask_hw[1] + 31) // 32 + 2
        mask_bits_bytes = n_img * (w.mask_hw[0] + 2) * pitch * 4
    alg_bytes = B * (w.vol_bytes() + w.feats_bytes()) + mask_bits_bytes
    achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch()
    roofline = {"bound": "hbm", "kernel": "sweep_kernel", "achieved": achieved, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if isinstance(traffic, dict) else None,
                "traffic_config": traffic.get("config") if isinstance(traffic, dict) else None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": sweep_ms,
                "write_only_frac": (B * w.vol_bytes() / (sweep_ms * 1e-3) / 1e9) / peak}
    if variants:
        for v in variants.values():
            v["frac"] = alg_bytes / (v["kernel_ms"] * 1e-3) / 1e9 / peak
        roofline["variants"] = variants
    # whole-step algorithmic bytes incl. the reference-interface float masks and the soft-argmin
    step_bytes = B * (w.vol_bytes() + w.feats_bytes() + w.masks_bytes(args.masks == "shared") +
                      w.softargmin_bytes(args.norm_costs))
    step_gbs = step_bytes / (ms_total_max / K * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, ms, cores = cpu_reference_throughput(w, 2, 3, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"2 frames of {w.name} per iteration, 1 warm-up + 3 timed, torch-eager port of the reference "
                         f"path (oracle/ref_port_torch.py), {cores} threads"}

    line = {
        "metric": "sweep+cost-volume+softargmin frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, w),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "latency_single_frame": latency, "step_algorithmic_gbs": step_gbs, "step_frac_of_peak": step_gbs / peak, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from mvs_gi_b200 import synthetic

    w = synthetic.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
