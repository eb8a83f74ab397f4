#!/bin/bash
# Multi-GPU pass (on a box with N GPUs): frame-sharded bench line at N GPUs + candidate-sharded sweep check/timing.
# Usage: bash profiles/runners/gpu_multi.sh <tag> <N>
tag=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$tag.log 2>&1; tail -1 gpurun_out/bench_n${N}_$tag.log
timeout 600 $TR scripts/dshard_bench.py --workload HIGHD --batch 1 --steps 10 > gpurun_out/dshard_n${N}_$tag.log 2>&1; tail -1 gpurun_out/dshard_n${N}_$tag.log
timeout 600 $TR scripts/dshard_bench.py --workload G16V --batch 16 --steps 10 > gpurun_out/dshard_g16v_n${N}_$tag.log 2>&1; tail -1 gpurun_out/dshard_g16v_n${N}_$tag.log
