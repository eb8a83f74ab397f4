#!/bin/bash
# round 2, first run of the tiled kernel: bit-identity tests, then both kernels through bench.py
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -15
for k in cta tiled; do
  for m in per_frame shared; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --kernel $k --masks $m > gpurun_out/r2a_${k}_${m}.json 2> gpurun_out/r2a_${k}_${m}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2a_${k}_${m}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$k $m", "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3))
except Exception as e:
    print("$k $m failed", e, open("gpurun_out/r2a_${k}_${m}.err").read()[-2000:])
PY
  done
done
