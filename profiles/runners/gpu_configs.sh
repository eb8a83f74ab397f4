#!/bin/bash
# Kernel times of every BASELINE.json configuration (parity cases, not bench lines).  Usage: bash profiles/runners/gpu_configs.sh <tag>
tag=${1:-cfg}
mkdir -p gpurun_out; : > gpurun_out/configs_$tag.jsonl
run() {  # workload batch extra...
  wl=$1; b=$2; shift 2
  timeout 300 python bench.py --workload $wl --batch $b --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-variants "$@" 2>gpurun_out/configs_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'workload':'$wl','batch':$b,'args':'$*','frames_per_s':d['value'],'ms_per_step':d['ms_per_step'],'sweep_ms':r['kernel_ms'],'sweep_frac':r['frac'],'sweep_write_only_frac':r['write_only_frac'],'alg_bytes':r['algorithmic_bytes_per_launch']}))" | tee -a gpurun_out/configs_$tag.jsonl || tail -3 gpurun_out/configs_err.log
}
run G16V 32
run G16V 32 --masks shared
run G16V 32 --feats-layout nhwc
run G16V 1
run G8 32
run E8 32
run G16VV_C48 16
run 4CAM 32
run G16V_DS 32
run HIGHD 2
