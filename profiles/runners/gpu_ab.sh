#!/bin/bash
# A/B of an experiment build against the default library.  Usage: bash profiles/runners/gpu_ab.sh <exp-name> <tag>
exp=$1; tag=${2:-ab}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$exp.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for lib in default $exp; do
  if [ $lib = default ]; then unset MVSGI_B200_LIB; else export MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_$exp.so; fi
  for i in 1 2; do
    timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/bench_${tag}_$lib.log 2>&1
    tail -1 gpurun_out/bench_${tag}_$lib.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))" || tail -5 gpurun_out/bench_${tag}_$lib.log
  done
done
unset MVSGI_B200_LIB
python scripts/backward_bench.py 16 | tee gpurun_out/backward_bench_$tag.json
