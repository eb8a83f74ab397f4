#!/bin/bash
# last check of the final tree: smoke (now with the kept-table path) and the GPU suite
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r4l_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r4l_pytest_gpu.log
