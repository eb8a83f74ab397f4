#!/bin/bash
# Tuning sweep over the planner's environment knobs (GPU box).  Prints sweep-kernel ms per configuration.
run() { echo -n "$* : "; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sweep_ms', round(d['roofline']['kernel_ms'],3), 'step_ms', round(d['ms_per_step'],3))"; }
for cfg in "$@"; do run $cfg; done
