#!/bin/bash
# round 2, tile-walking kernel: bit-identity tests with the shipped library, then the CTA-shape / frames-per-pass /
# tile-height sweep over the experiment libraries (bench shape only), static masks, sweep kernel time from bench.py
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -8
one() {  # lib fb tile_h layout masks
  local lib=$1 fb=$2 th=$3 lay=$4 m=$5
  local out=gpurun_out/${tag}_${lib}_fb${fb}_th${th}_${lay}_${m}.json
  MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_${lib}.so MVSGI_FRAMES_PER_PASS=$fb MVSGI_TILE_H=$th \
    timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --masks $m --feats-layout $lay > $out 2> ${out%.json}.err
  python - "$out" "$lib fb$fb th$th $lay $m" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), flush=True)
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1][:-5] + ".err").read()[-600:], flush=True)
PY
}
for lib in w16c2 w16c1 w8c3 w8c2 w8c4 w12c2 w10c2 w4c4; do
  for fb in 1 2; do
    one $lib $fb 8 nchw shared
  done
done
for lib in w16c1 w8c2 w8c3 w16c2; do
  one $lib 2 4 nchw shared
  one $lib 2 2 nchw shared
done
one w16c1 2 8 nhwc shared
one w8c2 2 8 nhwc shared
one w16c1 2 8 nchw per_frame
one w8c2 2 8 nchw per_frame
