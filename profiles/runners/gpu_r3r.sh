#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
timeout 300 python scripts/backward_bench.py 16 | tee gpurun_out/r3r_backward_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
