#!/bin/bash
# 2 GPUs: the candidate-sharded tests, then the bench line with the dshard block (kept tap tables in the fused path)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or sharded" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r4d_bench_2gpu.json 2> gpurun_out/r4d_bench_2gpu.err; tail -c 1500 gpurun_out/r4d_bench_2gpu.json; tail -3 gpurun_out/r4d_bench_2gpu.err
