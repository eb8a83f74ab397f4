#!/bin/bash
# N-GPU bench record: bash profiles/runners/gpu_r4i.sh N
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r4i_bench_${N}gpu.json 2> gpurun_out/r4i_bench_${N}gpu.err; tail -c 700 gpurun_out/r4i_bench_${N}gpu.json; tail -2 gpurun_out/r4i_bench_${N}gpu.err
