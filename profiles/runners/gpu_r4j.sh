#!/bin/bash
mkdir -p gpurun_out
timeout 500 env PYTHONPATH=$PWD compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/runners/sanitize_r4.py > gpurun_out/r4j_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r4j_memcheck.log
timeout 500 env PYTHONPATH=$PWD compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/runners/sanitize_r4.py > gpurun_out/r4j_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r4j_racecheck.log
timeout 300 env PYTHONPATH=$PWD compute-sanitizer --tool synccheck --error-exitcode 7 python profiles/runners/sanitize_r4.py > gpurun_out/r4j_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r4j_synccheck.log
