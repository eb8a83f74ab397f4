#!/bin/bash
run() {  # workload batch env...
  wl=$1; b=$2; shift 2
  env "$@" python bench.py --workload $wl --batch $b --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl $*', round(d['value']), round(r['kernel_ms'],4))"
}
run G16V 32 MVSGI_GROUP_PX=44
run G16V 32 MVSGI_GROUP_PX=100000
for wl in "G8 32" "4CAM 32" "HIGHD 2" "G16VV_C48 16" "G16V_DS 32"; do
  run $wl MVSGI_GROUP_PX=8
  run $wl MVSGI_GROUP_PX=100000
done
MVSGI_GROUP_PX=100000 python bench.py --vol-layout ncdhw --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-variants 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ncdhw big group', round(d['value']), round(r['kernel_ms'],4))"
