#!/bin/bash
tag=${1:-pairs}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pv in 0 1; do
  for i in 1 2; do
    MVSGI_PAIRS=$pv timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/bench_${tag}_$pv.log 2>&1
    tail -1 gpurun_out/bench_${tag}_$pv.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pairs=$pv fps',round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'sweep_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))" || tail -5 gpurun_out/bench_${tag}_$pv.log
  done
done
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"sweep_kernel|stage_feats" -c 12 --csv --log-file gpurun_out/ncu_${tag}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_${tag}.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
ids=sorted({int(r[ii]) for r in rows[1:]})
for r in rows[1:]:
    if int(r[ii]) in ids[-2:]: print(r[ii], r[ki][12:50], r[mi], r[vi])
PY
