#!/bin/bash
# N-GPU pass: frame-sharded bench (weak scaling), candidate-sharded sweep in both volume layouts, 2-GPU parity test.
# Usage (on an N-GPU box): bash profiles/runners/gpu_multi2.sh <N> <tag>
N=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_gpus 2>&1 | tail -2
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_n$N.log 2>&1; tail -1 gpurun_out/bench_${tag}_n$N.log | cut -c1-1200
for wl in HIGHD G16V; do
  b=1; [ $wl = G16V ] && b=16
  for cl in "" "--channels-last"; do
    timeout 600 $TR --master-port 29542 scripts/dshard_bench.py --workload $wl --batch $b --steps 20 $cl 2>/dev/null | grep '^{' | tee -a gpurun_out/dshard_${tag}_n$N.jsonl
  done
done
