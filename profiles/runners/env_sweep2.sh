#!/bin/bash
# Planner knobs of the sweep kernel (env overrides read by validate_and_plan) on the bench default.
run() {
  env "$@" python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*', round(d['value']), round(r['kernel_ms'],4))"
}
run MVSGI_INTERLEAVE=1
run MVSGI_INTERLEAVE=0 MVSGI_GROUP_PX=100000
run MVSGI_DCH=8
run MVSGI_TABLE_BYTES=98304 MVSGI_CTAS_PER_SM=2
