#!/usr/bin/env python
"""Text summary of an .ncu-rep capture (raw page): the metrics the profiles/ notes quote.  usage: ncu_summary.py <rep>"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for n, r in enumerate(rows[2:]):
    name = r[H.index("Kernel Name")] if "Kernel Name" in H else "?"
    print(f"## launch {n}: {name[:90]}")
    for k in KEYS:
        if k in H and r[H.index(k)] != "":
            print(f"  {k:75s} {r[H.index(k)]:>16s} {U[H.index(k)]}")
    st = [(float(r[i] or 0), h) for i, h in enumerate(H) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    print("  warp-cycles per issued instruction by stall reason (smsp__average_warps_issue_stalled_*_per_issue_active):")
    for v, h in sorted(st, reverse=True)[:9]:
        print(f"    {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
