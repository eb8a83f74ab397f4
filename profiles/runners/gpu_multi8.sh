#!/bin/bash
# 8-GPU confirmation: frame-sharded bench and the fused candidate-sharded sweep (HIGHD) in both layouts.
N=${1:-8}; tag=${2:-r1i}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/bench_${tag}_n$N.log 2>&1; tail -1 gpurun_out/bench_${tag}_n$N.log | cut -c1-400
for cl in "" "--channels-last"; do
  timeout 400 $TR --master-port 29552 scripts/dshard_bench.py --workload HIGHD --batch 1 --steps 20 $cl 2>/dev/null | grep '^{' | tee -a gpurun_out/dshard_${tag}_n$N.jsonl
done
