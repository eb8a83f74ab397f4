#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
one() {  # workload batch layout masks
  local w=$1 b=$2 lay=$3 m=$4
  local out=gpurun_out/r2h_${w}_b${b}_${lay}_${m}.json
  timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --workload $w --batch $b --masks $m --feats-layout $lay > $out 2> ${out%.json}.err
  python - "$out" "$w b$b $lay $m" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), flush=True)
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1][:-5] + ".err").read()[-600:], flush=True)
PY
}
for lay in nchw nhwc; do
  one G16V 32 $lay shared
  one G16V 32 $lay per_frame
done
one 4CAM 32 nhwc per_frame
one G16VV_C48 16 nhwc per_frame
one G8 32 nhwc per_frame
one HIGHD 2 nhwc per_frame
one G16V 1 nhwc shared
