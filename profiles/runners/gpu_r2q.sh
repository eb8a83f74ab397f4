#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inplace.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
one() {  # unit dch label [extra env]
  local out=gpurun_out/r2q_u$1_d$2_$3.json
  MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_base.so MVSGI_UNIT_OCT=$1 MVSGI_DCH=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --workload $3 --batch $4 > $out 2> ${out%.json}.err
  python - "$out" "unit $1 dch $2 $3" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), flush=True)
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1][:-5] + ".err").read()[-600:], flush=True)
PY
}
one 1 16 G16V 32
one 2 16 G16V 32
one 2 8 G16V 32
one 1 8 G16V 32
one 4 8 G16V 32
one 4 4 G16V 32
one 2 4 G16V 32
one 1 16 4CAM 32
one 2 8 4CAM 32
one 2 16 4CAM 32
one 1 16 G8 32
one 2 8 G8 32
