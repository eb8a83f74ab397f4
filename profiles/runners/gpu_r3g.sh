#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resample.py -m gpu -x -q 2>&1 | tail -5
