#!/bin/bash
# tests with the shipped library, then ncu of the tile-walking kernel in two CTA shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -8
for lib in w8c4 w16c2; do
  export MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_${lib}.so
  bash profiles/runners/gpu_ncu2.sh r2d_$lib sweep_tiled_kernel --masks shared | grep -E "duration|hit_rate|wavefronts|inst_executed|issue_active|tex_op_read|warps_active|stalled|registers|xbar2l1"
done
