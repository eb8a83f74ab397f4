#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep, read here without a GPU) into the short text summary committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_sweep_<tag>.ncu-rep profiles/r1_sweep_<tag>.txt [--json profiles/sweep_traffic.json --config "..."]
"""
import argparse
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--json")
    ap.add_argument("--config", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {a.rep}", f"# {a.config}"]
    got = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"kernel: {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"  {k:88s} {units[i]:16s} {r[i]}")
                got[k] = (r[i], units[i])
        def num(k):
            v, u = got[k]
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        try:
            ld_s, ld_r = num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
            lines.append(f"  derived: global-load sectors/request = {ld_s / max(ld_r, 1):.2f}; "
                         f"dram read+write = {(num('dram__bytes_read.sum') + num('dram__bytes_write.sum')) / 1e6:.1f} MB")
        except Exception as e:  # noqa: BLE001
            lines.append(f"  derived: n/a ({e})")
        if a.json:
            traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            with open(a.json, "w") as f:
                json.dump({"kernel": name[:80], "dram_bytes_per_launch": traffic,
                           "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
                           "duration_us_under_ncu": float(got["gpu__time_duration.sum"][0].replace(",", "")),
                           "source": a.rep, "config": a.config}, f, indent=1)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    sys.exit(main())
