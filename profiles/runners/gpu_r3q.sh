#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -x -q -k "backward or adjoint or whole_model or grad" 2>&1 | tail -12
timeout 300 python scripts/backward_bench.py 16 | tee gpurun_out/r3q_backward_bench.json
