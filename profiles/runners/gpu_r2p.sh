#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python scripts/postvol_layout.py G16V 8 > gpurun_out/r2p_postvol.json 2> gpurun_out/r2p_postvol.err; tail -c 300 gpurun_out/r2p_postvol.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2p_postvol.json").read().strip().splitlines()[-1])
for k, v in d.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()} if isinstance(v, dict) else v)
PY
# bf16 volume sweep kernel time at the bench shape
python - <<'PY'
import torch
from mvs_gi_b200 import abi, ops, synthetic
w = synthetic.WORKLOADS["G16V"]; dev = torch.device("cuda:0"); B = 32
rig_blob = ops.rig_to_blob(w.rig().to_struct())
masks = ops.pack_masks(torch.from_numpy(synthetic.fisheye_footprint_masks(w)).to(dev)[None])
sets = [torch.randn((B, 3, 128, 512, 16), device=dev).permute(0, 1, 4, 2, 3) for _ in range(3)]
for bf in (False, True):
    calls = [ops.SweepCall(f, None, None, masks, rig_blob, 0, 80, 320, 16, channels_last=True, mask_hw=w.mask_hw, vol_bf16=bf) for f in sets]
    for i in range(6): calls[i % 3].run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30): calls[i % 3].run()
    b.record(); torch.cuda.synchronize()
    print("vol_bf16", bf, "sweep ms per 32 frames", round(a.elapsed_time(b) / 30, 4))
PY
