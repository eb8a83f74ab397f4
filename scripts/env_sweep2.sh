#!/bin/bash
# Planner knobs of the sweep kernel (env overrides read by validate_and_plan) on the bench default.
run() {
  env "$@" python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*', round(d['value']), round(r['kernel_ms'],4))"
}
run MVSGI_GROUP_PX=8
run MVSGI_GROUP_PX=16
run MVSGI_GROUP_PX=32
run MVSGI_GROUP_PX=88
run MVSGI_DCH=8
run MVSGI_DCH=4
run MVSGI_INTERLEAVE=0 MVSGI_GROUP_PX=32
run MVSGI_INTERLEAVE=0 MVSGI_GROUP_PX=8
run MVSGI_CTAS_PER_SM=1
run MVSGI_TABLE_BYTES=65536
