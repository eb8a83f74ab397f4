#!/bin/bash
tag=${1:-f4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python scripts/resample_bench.py 24 > gpurun_out/resample_bench_$tag.json 2> gpurun_out/resample_bench_$tag.err; cat gpurun_out/resample_bench_$tag.json; tail -3 gpurun_out/resample_bench_$tag.err
timeout 300 python scripts/resample_bench.py 3 | tee gpurun_out/resample_bench_${tag}_n3.json
