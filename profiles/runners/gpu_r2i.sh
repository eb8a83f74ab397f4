#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 600 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), "launches", d["gpu_launches"])
for k, v in (r.get("variants") or {}).items():
    print("  ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "tolerance"})
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("numa",)})
print("numa", d["e2e"]["numa"])
print("latency", d["latency_single_frame"])
print("cpu", d["cpu_baseline"]["value"], "ref_on_gpu", d["reference_on_gpu"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
timeout 600 python bench.py --impl reference --device cuda --steps 5 --warmup 2 | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --masks per_frame --feats-layout nchw --no-cpu-baseline --no-variants | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference-interface config: value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['copy_ceiling']['value']))"
