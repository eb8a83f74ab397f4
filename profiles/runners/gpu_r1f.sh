#!/bin/bash
# GPU pass r1f: parity tests (incl. channels-last volume), bench in both volume layouts, predicated-tap experiment,
# post_vol layout timing.  Usage (on the GPU box): bash profiles/runners/gpu_r1f.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_r1f.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1f.log
for lay in ncdhw ndhwc; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --vol-layout $lay > gpurun_out/bench_r1f_$lay.log 2>&1
  tail -1 gpurun_out/bench_r1f_$lay.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lay fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'other', r['other_vol_layout'])"
done
for lay in ncdhw ndhwc; do
  MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_pred.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --vol-layout $lay > gpurun_out/bench_r1f_pred_$lay.log 2>&1
  tail -1 gpurun_out/bench_r1f_pred_$lay.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pred $lay fps',round(d['value']), 'sweep_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3))"
done
timeout 300 python scripts/postvol_layout.py G16V 8 > gpurun_out/postvol_layout_r1f.json 2> gpurun_out/postvol_layout_r1f.err; cat gpurun_out/postvol_layout_r1f.json; tail -2 gpurun_out/postvol_layout_r1f.err
