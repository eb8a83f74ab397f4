#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r3p_bench_2gpu.json 2> gpurun_out/r3p_bench_2gpu.err; tail -c 1200 gpurun_out/r3p_bench_2gpu.json; tail -3 gpurun_out/r3p_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -c 400
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or sharded" 2>&1 | tail -3
