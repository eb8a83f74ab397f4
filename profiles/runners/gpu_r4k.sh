#!/bin/bash
# The same bench with nothing precomputed outside the timed region (per-call projection, round-robin octets), and with
# the balanced deal only — the headline's switches taken off one at a time.
mkdir -p gpurun_out
timeout 600 python bench.py --tap-table off --balance off --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r4k_bench_nothing_kept.json 2> gpurun_out/r4k_a.err; tail -c 400 gpurun_out/r4k_bench_nothing_kept.json
timeout 600 python bench.py --tap-table off --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r4k_bench_balanced_only.json 2> gpurun_out/r4k_b.err; tail -c 400 gpurun_out/r4k_bench_balanced_only.json
