#!/bin/bash
tag=${1:-diag}
mkdir -p gpurun_out
bash scripts/microbench/run_all.sh $tag > /dev/null 2>&1; cat gpurun_out/microbench_$tag.log
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__f_wavefronts.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:sweep_kernel -c 8 --csv --log-file gpurun_out/ncu_${tag}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_${tag}.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
ids=sorted({int(r[ii]) for r in rows[1:]})
for r in rows[1:]:
    if int(r[ii]) in ids[-3:]: print(r[ii], r[ki][12:46], r[mi], r[vi])
PY
