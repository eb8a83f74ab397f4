#!/bin/bash
tag=${1:-bf16}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$tag.log 2>&1
tail -1 gpurun_out/bench_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), json.dumps(r.get('variants')))" || tail -5 gpurun_out/bench_$tag.log
