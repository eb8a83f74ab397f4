#!/bin/bash
tag=${1:-walk2}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --vol-layout ndhwc > gpurun_out/bench_${tag}_$name.log 2>&1
  tail -1 gpurun_out/bench_${tag}_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'other', r['other_vol_layout']['kernel_ms'])" || tail -5 gpurun_out/bench_${tag}_$name.log
}
run pf0 MVSGI_WALK_SEGS=16
for pf in 1 2 3 5 8; do
  run pf${pf} MVSGI_WALK_SEGS=16 MVSGI_WALK_PF=$pf
  run pf${pf}_2 MVSGI_WALK_SEGS=16 MVSGI_WALK_PF=$pf MVSGI_WALK_PF2=1
done
run pf3_segs8 MVSGI_WALK_SEGS=8 MVSGI_WALK_PF=3
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum --clock-control none -k regex:sweep -c 6 --csv --log-file gpurun_out/ncu_${tag}_pf0.csv env MVSGI_WALK_SEGS=16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --vol-layout ndhwc > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum --clock-control none -k regex:sweep -c 6 --csv --log-file gpurun_out/ncu_${tag}_pf3.csv env MVSGI_WALK_SEGS=16 MVSGI_WALK_PF=3 MVSGI_WALK_PF2=1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --vol-layout ndhwc > /dev/null 2>&1
python - <<PY
import csv
for f in ("gpurun_out/ncu_${tag}_pf0.csv","gpurun_out/ncu_${tag}_pf3.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
    for r in rows[1:]:
        if r[ii] in ('4','5'): print(f.split('_')[-1], r[ii], r[ki][:30], r[mi], r[vi])
PY
