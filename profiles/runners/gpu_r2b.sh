#!/bin/bash
# tiled-kernel iteration: bit-identity tests, bench (both mask modes), ncu summary of the tiled kernel
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -5
for m in per_frame shared; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --masks $m > gpurun_out/${tag}_tiled_${m}.json 2> gpurun_out/${tag}_tiled_${m}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_tiled_${m}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("tiled $m", "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3))
except Exception as e:
    print("tiled $m failed", e, open("gpurun_out/${tag}_tiled_${m}.err").read()[-2000:])
PY
done
bash profiles/runners/gpu_ncu2.sh ${tag}_tiled sweep_tiled_kernel --masks shared | grep -E "duration|hit_rate|wavefronts|inst_executed|issue_active|tex_op_read|warps_active|stalled_long|registers"
