#!/bin/bash
# Build and run every microbenchmark on the GPU box; output -> gpurun_out/microbench_<tag>.log
tag=${1:-x}
cd "$(dirname "$0")"
for f in l1_taps lsu_cost l2_chunks st_shapes; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $f $f.cu || exit 1
  echo "== $f"; timeout 120 ./$f
done > ../../gpurun_out/microbench_$tag.log 2>&1
cat ../../gpurun_out/microbench_$tag.log
