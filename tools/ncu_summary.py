#!/usr/bin/env python
"""Summary of an `ncu --set full` report: one record per captured kernel launch with the counters DESIGN.md / bench.py
quote (profiles/r2_frame_ncu_summary.json is made by this script; bench.py reads `dram__bytes_*` and
`smsp__inst_executed.sum` from it).  Usage: tools/ncu_summary.py <report.ncu-rep> [out.json]"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "lts__t_sectors.sum",
    "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum",
    "lts__t_sectors_op_atom.sum",
    "lts__t_sectors_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        rec = {"Kernel Name": d["Kernel Name"]}
        for k in KEEP:
            if k in d and d[k] != "":
                rec[k] = ("%s %s" % (d[k].replace(",", ""), u[k])).strip()
        if "lts__t_sectors.sum" in rec:
            rec["lts__t_bytes (sectors x 32 B)"] = "%.1f Mbyte" % (float(rec["lts__t_sectors.sum"].split()[0]) * 32 / 1e6)
        recs.append(rec)
    text = json.dumps(recs, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
