#!/bin/bash
# usage: bash profiles/runners/gpu_multi_r2.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 800 gpurun_out/r2_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4))
e = d["e2e"]
print("e2e", round(e["value"]), "ceiling", round(e["copy_ceiling"]["value"]), "frac", round(e["frac_of_copy_ceiling"], 3), "h2d GB/s all ranks", round(e["copy_ceiling"]["h2d_gbs_all_ranks"], 1), e["numa"])
print("dshard", json.dumps(d["dshard"]))
PY
