#!/usr/bin/env python3
"""Summarises an .ncu-rep (ncu --set full) as a markdown table: one column per captured launch.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rN_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "sm__icc_request_hit_rate.pct", "gcc__average_cache_request_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    if rep.endswith(".csv"):        # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): reports of the
        txt = open(rep).read()      # Atlas library exceed the 64 MiB that gpurun copies back
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print("# %s\n" % title)
    print("Source: `ncu --set full --clock-control none --import-source on` (per-launch times under ncu are "
          "cold-cache and serialised: compare shares and ratios, not absolutes).\n")
    print("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join("`%s`" % r[name_i].split("(")[0].replace("void ", "") for r in data) + " |")
    for k in KEYS:
        if k not in hdr:
            continue
        i = hdr.index(k)
        print("| %s | %s | %s |" % (k, units[i], " | ".join(r[i] for r in data)))


if __name__ == "__main__":
    main()
