"""Markdown summary of one kernel's ncu --set full capture: usage ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, val = rows[0], rows[1], rows[2]
d = dict(zip(hdr, val))
u = dict(zip(hdr, units))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in d:
        print("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
print("\nWarp stall reasons (stalled warps per issue-active cycle):")
for h in hdr:
    if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        print("* %s: %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[h]))
