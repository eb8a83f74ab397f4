#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -8
one() {  # workload batch kernel layout masks
  local w=$1 b=$2 k=$3 lay=$4 m=$5
  local out=gpurun_out/r2f_${w}_b${b}_${k}_${lay}_${m}.json
  timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --workload $w --batch $b --kernel $k --masks $m --feats-layout $lay > $out 2> ${out%.json}.err
  python - "$out" "$w b$b $k $lay $m" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), flush=True)
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1][:-5] + ".err").read()[-600:], flush=True)
PY
}
for k in cta tiled; do
  for lay in nchw nhwc; do
    one G16V 32 $k $lay shared
    one G16V 32 $k $lay per_frame
  done
  one 4CAM 32 $k nhwc per_frame
  one HIGHD 2 $k nhwc per_frame
  one G16VV_C48 16 $k nhwc per_frame
  one G8 32 $k nhwc per_frame
  one G16V_DS 32 $k nhwc per_frame
  one G16V 1 $k nhwc shared
done
