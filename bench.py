#!/usr/bin/env python
"""bench.py — sweep + cost-volume + soft-argmin frames/s on N B200 (G16V shape) and % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--device cpu|cuda]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of B synthetic frames per GPU:
    fused sweep/aggregate kernel -> soft-argmin kernel
(plus, only when the inputs ask for them: NCHW -> channels-last staging, mask bit-packing).
Defaults are the inference configuration north_star describes: features stored channels-last (tapped in place), the
rig's masks static and resident on the device (packed once, as api/inference_pytorch.py keeps them), in-kernel
projection, channels-last volume.  The reference-interface variants (NCHW features from cuDNN, per-frame float masks,
NCDHW volume) are timed in the same run and reported under `roofline.variants`.
Frames are independent, so N GPUs shard frames with no collective (weak scaling: B frames per GPU).
`value`      : whole-job frames/s with inputs resident in HBM (device-timed, max over ranks).
`e2e`        : same metric through the public module API with HOST (pinned) inputs: H2D of feats+costs and D2H of
               inv_dist inside the timed region; `copy_ceiling` = the same copies without any kernel.
`roofline`   : the dominant kernel (sweep_kernel) timed on its own with CUDA events inside the timed region.
`cpu_baseline`: the PyTorch-eager port of the reference path (oracle/ref_port_torch.py) on this box's host cores.
`reference_on_gpu`: the same port with its tensors on the B200 (how api/inference_pytorch.py:115-122 runs the path).
`--impl reference`: the reference arm (the port on the host cores; `--device cuda` puts it on the GPU).
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
METRIC = "sweep+cost-volume+softargmin frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: where the port runs")
    ap.add_argument("--workload", default="G16V")
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--masks", default="static", choices=["static", "per_frame", "shared"],
                    help="static = the rig's masks, resident on the device and packed once (inference); per_frame = the "
                         "reference interface, [B,CAM,1,Hm,Wm] float32 masks read and packed on every call (training "
                         "with RandomSingleImageMasking); shared = alias of static")
    ap.add_argument("--feats-layout", default="nhwc", choices=["nchw", "nhwc"],
                    help="memory layout the feature extractor hands over: nhwc = channels-last (north_star (a), tapped "
                         "in place); nchw = cuDNN's default, staged by an extra pass")
    ap.add_argument("--vol-layout", default="ndhwc", choices=["ncdhw", "ndhwc"],
                    help="memory layout of the cost volume: ncdhw = the reference's; ndhwc = channels_last_3d strides "
                         "(MVSGI_FLAG_VOL_CHANNELS_LAST, what cuDNN's NDHWC Conv3d of post_vol consumes)")
    ap.add_argument("--balance", default="on", choices=["on", "off"],
                    help="static masks: deal the output octets to the sweep's CTAs by their precomputed cost "
                         "(ops.balanced_octet_order); off = round-robin octets")
    ap.add_argument("--tap-table", default="on", choices=["on", "off"],
                    help="static masks: keep the sweep kernel's tap tables across calls (ops.build_tap_table: built "
                         "once outside the timed region — the role of the reference's precomputed grids — and copied "
                         "in by cp.async.bulk at every call); off = every call evaluates the geometry")
    ap.add_argument("--norm-costs", action="store_true", help="also write norm_costs (training consumers)")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps, for `ncu --profile-from-start off`: the "
                         "capture then holds the timed launches only (no calibration / table-build launches)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the other-layout / bf16-staging kernel timings")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline and reference_on_gpu")
    ap.add_argument("--no-dshard", action="store_true", help="N > 1: skip the candidate-sharded comparison block")
    ap.add_argument("--ref-frames", type=int, default=1, help="frames per step of the reference arm")
    args = ap.parse_args()
    if args.masks == "shared":
        args.masks = "static"
    return args


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


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML lists as local to GPU `index` (its NUMA node) BEFORE pinned host buffers are
    allocated, so that they live in the memory the GPU's PCIe root is attached to.  Returns what was done."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return {"bound": True, "cpus": f"{allowed[0]}-{allowed[-1]}", "n": len(allowed)}
        return {"bound": False, "cpus": f"{cpus[0]}-{cpus[-1]}" if cpus else None,
                "why": "the GPU's local CPU set is the whole allowed set (one NUMA node)"}
    except Exception as e:  # NVML missing or no permission: not fatal
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


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


# ------------------------------------------------------------------------------------------- reference arm (the port)
def reference_throughput(w, frames: int, iters: int, warmup: int, device: str = "cpu"):
    """frames/s of the PyTorch-eager port of the reference path: on the host cores (all threads), or with its tensors
    on the GPU (`device="cuda"`, how the reference's own inference API runs it; TF32 is irrelevant, the path has no
    matmul or convolution)."""
    import numpy as np
    import torch

    from mvs_gi_b200 import synthetic
    from oracle import oracle as orc
    from oracle import ref_port_torch as port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device(device)
    rs = w.rig().to_struct()
    g, gm = orc.make_grids(rs)  # the grids the reference would have precomputed on the CPU
    grids = torch.from_numpy(g)[None].repeat(frames, 1, 1, 1, 1, 1).to(dev)
    gmask = torch.from_numpy(gm)[None, ..., None].repeat(frames, 1, 1, 1, 1, 1).to(dev)
    feats = torch.from_numpy(synthetic.make_feats(w, frames, 1234)).to(dev)
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w))[None].repeat(frames, 1, 1, 1, 1).to(dev)
    costs = torch.from_numpy(synthetic.make_costs(w, frames, 1234)).to(dev)
    idx = torch.tensor([w.bf / d for d in np.array(w.dist_list, dtype=np.float32)], dtype=torch.float32, device=dev)

    def step():
        with torch.no_grad():
            if w.aggregate == "std":
                port.sweep_std(feats, grids, gmask, masks)
            else:
                port.sweep_concat(feats, grids)
            port.soft_argmin(costs, idx, w.scale, True)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    sync()
    dt = time.perf_counter() - t0
    return frames * iters / dt, dt / iters * 1e3, cores


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    on_gpu = args.device == "cuda"
    fps, ms, cores = reference_throughput(w, args.ref_frames, args.steps, max(args.warmup, 1), args.device)
    where = "on cuda:0 (eager launches)" if on_gpu else f"{cores} host threads"
    sample = f"{args.ref_frames} frame(s) of {w.name} per step, {args.steps} steps, torch-eager port, {where}"
    cfg = workload_config(args, w)
    # what this arm actually ran: the reference interface (supplied grids, per-frame float masks, NCHW, NCDHW volume)
    cfg.update({"frames_per_gpu_per_step": args.ref_frames, "global_batch": args.ref_frames,
                "grids": "supplied [B,CAM,D,Ho,Wo,2] float32 (the reference precomputes them on the CPU)",
                "masks": "per_frame", "feats_layout": "nchw", "vol_layout": "ncdhw", "norm_costs": True,
                "parallelism": "rank 0 only", "cache": "same inputs every step",
                "runs": "n/a (the reference formulation: a python loop over candidates of ATen ops)",
                "tap_table": "n/a (the reference reads its precomputed grids at every call)",
                "device": "cuda:0" if on_gpu else "cpu"})
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
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
        "runs": ("output octets dealt to the sweep's CTAs by precomputed cost, corrected once by measured per-CTA times "
                 "(static rig + masks; computed outside the timed region, see run_balance)")
                if (args.balance == "on" and args.masks != "per_frame") else "round-robin octets",
        "tap_table": ("kept across calls: the kernel's geometry / validity records are built once outside the timed "
                      "region, as the reference precomputes its grids, and copied in by every launch")
                     if (args.tap_table == "on" and args.masks != "per_frame") else "built by every launch",
    }


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, w):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    numa = bind_to_gpu_numa_node(local)  # before any pinned allocation
    from mvs_gi_b200 import abi, modules, ops, synthetic

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi.load_library()

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    rig = w.rig()
    rig_blob = ops.rig_to_blob(rig.to_struct())
    agg = abi.AGG_STD if w.aggregate == "std" else abi.AGG_CONCAT
    std = w.aggregate == "std"
    nsets = 3
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    base_masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)

    def nhwc(f):
        return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)

    def mask_inputs(mode):
        """(masks tensor for SweepCall, mask_hw): static = bits packed once here, outside every timed region."""
        if not std:
            return None, None
        if mode == "per_frame":
            return torch.from_numpy(synthetic.erase_rectangles(base_masks.cpu().numpy(), B, 77 + rank)).to(dev), None
        return ops.pack_masks(base_masks[None], check=True), tuple(w.mask_hw)

    masks, mask_hw = mask_inputs(args.masks)
    feat_sets = [torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen) for _ in range(nsets)]
    if args.feats_layout == "nhwc":
        feat_sets = [nhwc(f) for f in feat_sets]
    calls = [ops.SweepCall(f, None, None, masks, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands,
                           channels_last=args.vol_layout == "ndhwc", mask_hw=mask_hw) for f in feat_sets]
    # static calibration + static masks: the octets of the output image are dealt to the sweep's CTAs by their known
    # cost (ops.balanced_octet_order) — computed once here, outside every timed region, like the packed masks
    octet_order, balance_report = None, None
    if args.balance == "on" and (not std or args.masks != "per_frame"):
        octet_order, balance_report = ops.calibrated_octet_order(calls[0])
        for c in calls:
            c.set_octet_order(octet_order)
    # ... and the kernel's tap tables (geometry + validity records) are kept across calls, as the reference keeps its
    # precomputed grids: one BUILD launch here, every timed launch copies them in
    tap_table = None
    if args.tap_table == "on" and (not std or args.masks != "per_frame"):
        tap_table = ops.build_tap_table(calls[0])
        for c in calls[1:]:
            c.set_tap_table(tap_table)
    costs_sets = [torch.randn((B, 1, w.num_cands, *w.cv_hw), device=dev, generator=gen) for _ in range(nsets)]
    reg = modules.DistanceRegressorWithFixedCandidates(w.bf, w.dist_list, w.scale, True).to(dev)
    Ho, Wo, scale = reg._out_size(*w.cv_hw)
    stage_launches = calls[0].launch_count(abi.FLAG_STAGE_ONLY)

    def step(i, ev=None):
        c = calls[i % nsets]
        if stage_launches:
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
    if args.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(K):
        step(i, evs[i])
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    launches = ops.launch_counter() - launches0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    sweep_ms = sum(a.elapsed_time(b) for a, b in evs) / K
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    value = world * B * K / (ms_total_max * 1e-3)

    # ---- the same inputs through the other input / output layouts, mask modes and the bf16-staged stated-tolerance
    # mode: staging passes and sweep kernel timed on their own; each timed launch follows a pass over another 400 MB
    # input set (L2 evicted)
    med = lambda v: sorted(v)[len(v) // 2]

    def time_variant(feats_layout, mask_mode, vol_cl, feats_bf16=False, balanced=None, kept=None):
        f = feat_sets[0] if feats_layout == args.feats_layout else \
            (nhwc(feat_sets[0]) if feats_layout == "nhwc" else feat_sets[0].contiguous())
        m, mhw = (masks, mask_hw) if mask_mode == args.masks else mask_inputs(mask_mode)
        oc = ops.SweepCall(f, None, None, m, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands, channels_last=vol_cl,
                           feats_bf16=feats_bf16, mask_hw=mhw)
        if balanced is None:
            balanced = octet_order is not None
        if balanced and (not std or mask_mode != "per_frame"):
            oc.set_octet_order(ops.calibrated_octet_order(oc)[0])
        if kept is None:
            kept = tap_table is not None
        if kept and (not std or mask_mode != "per_frame"):
            ops.build_tap_table(oc)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(6)]
        for e in ev:
            calls[1].run(abi.FLAG_SWEEP_ONLY if not stage_launches else abi.FLAG_STAGE_ONLY)
            e[0].record()
            oc.run(abi.FLAG_STAGE_ONLY)
            e[1].record()
            oc.run(abi.FLAG_SWEEP_ONLY)
            e[2].record()
        torch.cuda.synchronize()
        return {"feats_layout": feats_layout, "masks": mask_mode, "vol_layout": "ndhwc" if vol_cl else "ncdhw",
                "runs": "cost-balanced" if oc.args.octet_order else "round-robin octets",
                "tap_table": "kept" if oc.args.tap_table_mode == abi.TAP_TABLE_USE else "built per call",
                "stage_launches": oc.launch_count(abi.FLAG_STAGE_ONLY),
                "stage_ms": med([e[0].elapsed_time(e[1]) for e in ev[1:]]),
                "kernel_ms": med([e[1].elapsed_time(e[2]) for e in ev[1:]])}

    variants = None
    if rank == 0 and not args.no_variants:
        cl = args.vol_layout == "ndhwc"
        other_lay = "nchw" if args.feats_layout == "nhwc" else "nhwc"
        other_masks = "per_frame" if args.masks == "static" else "static"
        variants = {"this_config": time_variant(args.feats_layout, args.masks, cl),
                    "other_feats_layout": time_variant(other_lay, args.masks, cl),
                    "other_vol_layout": time_variant(args.feats_layout, args.masks, not cl)}
        if std:
            variants["other_masks"] = time_variant(args.feats_layout, other_masks, cl)
            variants["reference_interface"] = time_variant("nchw", "per_frame", False)
        variants["other_runs"] = time_variant(args.feats_layout, args.masks, cl, balanced=octet_order is None)
        variants["other_tap_table"] = time_variant(args.feats_layout, args.masks, cl, kept=tap_table is None)
        variants["feats_bf16"] = dict(time_variant(args.feats_layout, args.masks, cl, True),
                                      tolerance="max|vol - fp32 vol| <= 1e-2 * max|vol| (tests/test_gpu_parity.py)")
    barrier()

    # ---- single-frame latency (the reference deploys at batch 1, api/inference_pytorch.py): eager launches and the same
    # kernels captured once in a CUDA graph (the library neither allocates nor synchronises)
    latency = None
    if rank == 0 and not args.no_variants:
        f1 = feat_sets[0][:1]
        f1 = nhwc(f1.contiguous()) if args.feats_layout == "nhwc" else f1.contiguous()
        m1 = masks[:1] if masks is not None else None
        c1 = ops.SweepCall(f1, None, None, m1, rig_blob, agg, w.cv_hw[0], w.cv_hw[1], w.num_cands,
                           channels_last=args.vol_layout == "ndhwc", mask_hw=mask_hw)
        if tap_table is not None:
            ops.build_tap_table(c1)
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
        latency = {"frames": 1, "ms_eager": eager_ms, "ms_cuda_graph": graph_ms, "launches": c1.launch_count() + 1}
        del graph
    barrier()

    # ---- end to end through the public module API, host buffers
    e2e = None
    if not args.no_e2e:
        builder_cls = modules.SphericalSweepStdMasked if std else modules.SphericalSweep
        builder = builder_cls(w.num_cams, w.out_chs, 3).to(dev).eval().attach_rig(rig)
        per_frame = std and args.masks == "per_frame"
        if std and not per_frame:
            builder.set_static_masks(base_masks[None])  # inference: masks uploaded once (api/inference_class.py:40-45)
        builder.set_vol_channels_last(args.vol_layout == "ndhwc")
        reg.return_norm_costs = bool(args.norm_costs)
        # the extractor's output as the host holds it: channels-last when that is the configured layout
        if args.feats_layout == "nhwc":
            h_feats = [torch.randn((B, w.num_cams, *w.feat_hw, w.feat_chs)).pin_memory().permute(0, 1, 4, 2, 3)
                       for _ in range(2)]
        else:
            h_feats = [torch.randn((B, w.num_cams, w.feat_chs, *w.feat_hw)).pin_memory() for _ in range(2)]
        h_costs = [torch.randn((B, 1, w.num_cands, *w.cv_hw)).pin_memory() for _ in range(2)]
        h_masks = [torch.from_numpy(synthetic.erase_rectangles(base_masks.cpu().numpy(), B, 5 + s)).pin_memory()
                   for s in range(2)] if per_frame else None
        h_out = torch.empty((B, 1, Ho, Wo)).pin_memory()

        # double-buffered upload: the H2D copy of step i+1 runs on a copy stream while step i computes; every step's
        # inputs still cross PCIe from pinned host memory inside the timed region and its result goes back to the host
        copy_stream = torch.cuda.Stream()
        d_feats = [torch.empty_like(h_feats[0], device=dev) for _ in range(2)]
        d_costs = [torch.empty_like(h_costs[0], device=dev) for _ in range(2)]
        d_masks = [torch.empty_like(h_masks[0], device=dev) for _ in range(2)] if per_frame else None
        d_inv = torch.zeros((B, 1, Ho, Wo), device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            s_ = i % 2
            copy_stream.wait_event(done[s_])  # the step that last read this device buffer has finished
            with torch.cuda.stream(copy_stream):
                d_feats[s_].copy_(h_feats[s_], non_blocking=True)
                d_costs[s_].copy_(h_costs[s_], non_blocking=True)
                if per_frame:
                    d_masks[s_].copy_(h_masks[s_], non_blocking=True)
                ready[s_].record(copy_stream)

        def e2e_run(n, compute=True):
            cur = torch.cuda.current_stream()
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                s_ = i % 2
                cur.wait_event(ready[s_])
                with torch.no_grad():
                    inv = d_inv
                    if compute:
                        f, c = d_feats[s_], d_costs[s_]
                        if std:
                            builder.sweep(f, None, None, d_masks[s_] if per_frame else None)
                        else:
                            builder.sweep(f, None)
                        inv, _ = reg(c)
                    h_out.copy_(inv, non_blocking=True)
                done[s_].record(cur)

        def timed_e2e(n, compute):
            e2e_run(3, compute)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            e2e_run(n, compute)
            s1.record()
            barrier()
            t2 = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            return world * B * n / (float(t2.item()) * 1e-3)

        Ke = K
        e2e_fps = timed_e2e(Ke, True)
        checksum = float(h_out.sum())
        ceiling_fps = timed_e2e(Ke, False)  # the same copies, all ranks at once, no kernel: what PCIe + host memory allow
        h2d = int(h_feats[0].numel() * 4 + h_costs[0].numel() * 4 + (h_masks[0].numel() * 4 if per_frame else 0))
        e2e = {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(h_out.numel() * 4), "steps": Ke, "result_checksum": checksum,
               "masks": args.masks,
               "copy_ceiling": {"value": ceiling_fps, "unit": "frames/s",
                                "h2d_gbs_all_ranks": ceiling_fps / B * h2d / 1e9,
                                "what": "identical pinned H2D/D2H copies on every rank at once, no kernel launched"},
               "frac_of_copy_ceiling": e2e_fps / ceiling_fps, "numa": numa}

    # ---- N > 1: the optional candidate-sharded variant (BASELINE.json config 5) next to the frame-sharded product path
    dshard = None
    if world > 1 and not args.no_dshard:
        try:
            dshard = dshard_block(torch, dist, dev, rank, world)
        except Exception as e:  # never lose the main line over the optional block
            dshard = {"error": f"{type(e).__name__}: {e}"}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sweep_kernel), algorithmic bytes per launch (SURVEY.md §8d / DESIGN.md)
    peak, peak_kind = measured_peak_gbs()
    mask_bits_bytes = 0
    if std:
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
    # whole-step algorithmic bytes incl. the reference-interface float masks (per_frame) and the soft-argmin
    step_bytes = B * (w.vol_bytes() + w.feats_bytes() + w.masks_bytes(args.masks == "static") +
                      w.softargmin_bytes(args.norm_costs))
    step_gbs = step_bytes / (ms_total_max / K * 1e-3) / 1e9

    cpu = ref_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, ms, cores = reference_throughput(w, 2, 3, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"2 frames of {w.name} per iteration, 1 warm-up + 3 timed, torch-eager port of the reference "
                         f"path (oracle/ref_port_torch.py), {cores} threads"}
        gfps, gms, _ = reference_throughput(w, 2, 5, 2, "cuda")
        ref_gpu = {"value": gfps, "unit": "frames/s", "ms_per_frame": gms / 2, "kind": "port",
                   "sample": f"2 frames of {w.name} per iteration, 2 warm-up + 5 timed, the same torch-eager port with "
                             f"its tensors on this B200 (supplied grids, per-frame float masks, NCDHW volume)",
                   "speedup_of_value": value / gfps}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, w),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "run_balance": balance_report,
        "tap_table": ({"kept": True, "bytes": int(tap_table.numel()),
                       "note": "built once outside the timed region (ops.build_tap_table), read by every launch"}
                      if tap_table is not None else {"kept": False}),
        "latency_single_frame": latency, "step_algorithmic_gbs": step_gbs, "step_frac_of_peak": step_gbs / peak,
        "cpu_baseline": cpu, "reference_on_gpu": ref_gpu, "dshard": dshard,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dshard_block(torch, dist, dev, rank, world, steps: int = 5):
    """Candidate-sharded sweep of the high-D volume (BASELINE.json config 5, 419 MB) on the ranks of this run: the NCCL
    all-gather baseline and the fused peer-store kernel against the un-sharded sweep on one GPU — timings (max over
    ranks), NVLink GB/s and the bit-identity of every rank's volume.  Not the product path (frames are sharded)."""
    from mvs_gi_b200 import modules, sharding, synthetic

    w = synthetic.WORKLOADS["HIGHD"]
    rig = w.rig()
    gen = torch.Generator(device=dev).manual_seed(4321)  # same features on every rank
    feats = torch.randn((1, w.num_cams, w.feat_chs, *w.feat_hw), device=dev, generator=gen)
    masks = torch.from_numpy(synthetic.fisheye_footprint_masks(w)[None]).to(dev)
    mk = lambda: modules.SphericalSweepStdMasked(w.num_cams, w.feat_chs, 3).to(dev).eval()
    full = mk().attach_rig(rig).set_vol_channels_last(True)
    nccl = sharding.CandidateShardedSweep(mk(), rig)
    fused = sharding.CandidateShardedSweep(mk(), rig, fused=True)
    for m in (nccl.builder, fused.builder):
        m.set_vol_channels_last(True)
    for m in (full, nccl.builder, fused.builder):  # the inference configuration: the rig's masks, packed once
        m.set_static_masks(masks)

    def timed(fn):
        for _ in range(2):
            out = fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    with torch.no_grad():
        v_full, ms_full = timed(lambda: full.sweep(feats, None, None, masks))
        v_nccl, ms_nccl = timed(lambda: nccl.sweep(feats, None, None, masks))
        v_fused, ms_fused = timed(lambda: fused.sweep(feats, None, None, masks))
    same = bool(torch.equal(v_nccl, v_full)) and bool(torch.equal(v_fused, v_full))
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    vol_bytes = v_full.numel() * 4
    recv = vol_bytes * (world - 1) / world
    if fused._peer_volume is not None:
        fused._peer_volume.close()
    return {"workload": "HIGHD: std, 3 cams, C=16, D=64, cv 160x640, 1 frame, channels-last volume", "n_gpus": world,
            "vol_mbytes": vol_bytes / 1e6, "bit_identical_to_unsharded_on_all_ranks": bool(ok.item()),
            "ms_unsharded_one_gpu": ms_full, "ms_nccl_allgather_total": ms_nccl, "ms_fused_peer_store_total": ms_fused,
            "nvlink_recv_gbs_per_rank_fused": recv / (ms_fused * 1e-3) / 1e9,
            "nvlink_recv_gbs_per_rank_nccl": recv / (ms_nccl * 1e-3) / 1e9, "nvlink_peak_gbs_per_direction": 900}


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
