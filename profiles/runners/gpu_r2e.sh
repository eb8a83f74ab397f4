#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiled.py -x -q 2>&1 | tail -8
# microbenchmark: wavefronts per tap request, L1-resident vs L2-resident
timeout 120 ./scripts/microbench/tap_miss > gpurun_out/mb_tap_miss.log 2>&1; cat gpurun_out/mb_tap_miss.log
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,sm__cycles_elapsed.max,l1tex__data_bank_reads.sum,l1tex__data_bank_writes.sum --clock-control none --csv --log-file gpurun_out/mb_tap_miss_ncu.csv ./scripts/microbench/tap_miss > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/mb_tap_miss_ncu.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
byid = {}
for r in rows[hdr + 1:]:
    d = dict(zip(cols, r))
    byid.setdefault(int(d["ID"]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
for i in sorted(byid):
    if i % 2 == 0: continue   # warm launches
    m = byid[i]
    req = m["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
    print("launch", i, "win/pat", i // 8, (i // 2) % 4, "wavefronts/req %.2f" % (m["l1tex__data_pipe_lsu_wavefronts.sum"] / req),
          "sectors/req %.2f" % (m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"] / req), "hit %.1f" % m["l1tex__t_sector_hit_rate.pct"],
          "clk/req/SM %.2f" % (m["sm__cycles_elapsed.max"] * 148 / req), "bank rd/req %.2f wr/req %.2f" % (m["l1tex__data_bank_reads.sum"] / req, m["l1tex__data_bank_writes.sum"] / req))
PY
one() {  # lib fb tile_h layout masks
  local lib=$1 fb=$2 th=$3 lay=$4 m=$5
  local out=gpurun_out/r2e_${lib}_fb${fb}_th${th}_${lay}_${m}.json
  MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_${lib}.so MVSGI_FRAMES_PER_PASS=$fb MVSGI_TILE_H=$th \
    timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-variants --no-cpu-baseline --masks $m --feats-layout $lay > $out 2> ${out%.json}.err
  python - "$out" "$lib fb$fb th$th $lay $m" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "step_ms", round(d["ms_per_step"], 4), "sweep_ms", round(r["kernel_ms"], 4), "frac", round(r["frac"], 3), flush=True)
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1][:-5] + ".err").read()[-600:], flush=True)
PY
}
one w8c3 1 8 nchw shared
one w8c3 2 8 nchw shared
one w16c2 1 8 nchw shared
one w16c1 2 8 nchw shared
one w8c2 2 8 nchw shared
one w8c2 1 8 nchw shared
export MVSGI_B200_LIB=$PWD/mvs_gi_b200/libmvsgi_b200_w8c3.so MVSGI_FRAMES_PER_PASS=1
bash profiles/runners/gpu_ncu2.sh r2e_w8c3_fb1 sweep_tiled_kernel --masks shared | grep -E "duration|hit_rate|wavefronts|inst_executed|issue_active|tex_op_read|warps_active|stalled_long|registers|xbar2l1|requests"
