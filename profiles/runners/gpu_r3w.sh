#!/bin/bash
mkdir -p gpurun_out
timeout 1200 env PYTHONPATH=$PWD compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/runners/sanitize_r3.py > gpurun_out/r3w_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r3w_memcheck.log
timeout 1200 env PYTHONPATH=$PWD compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/runners/sanitize_r3.py > gpurun_out/r3w_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r3w_racecheck.log
