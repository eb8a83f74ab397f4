#!/bin/bash
# GPU pass: walking-quad sweep (channels-last volume).  Parity tests, then the bench in ndhwc layout with the walker
# off / on and different segment counts.  Usage (on the GPU box): bash profiles/runners/gpu_walk.sh <tag>
tag=${1:-walk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -5 gpurun_out/pytest_gpu_$tag.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --vol-layout ndhwc > gpurun_out/bench_${tag}_$name.log 2>&1
  tail -1 gpurun_out/bench_${tag}_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'other', r['other_vol_layout']['kernel_ms'])" || tail -5 gpurun_out/bench_${tag}_$name.log
}
run walk0 MVSGI_WALK=0
run walk_default MVSGI_WALK=1
run walk_segs4 MVSGI_WALK_SEGS=4
run walk_segs16 MVSGI_WALK_SEGS=16
run walk_segs32 MVSGI_WALK_SEGS=32
run walk_dch8 MVSGI_DCH=8
run walk_shared_masks MVSGI_WALK=1
