"""Key counters of one kernel from an ncu report (raw page CSV): python tools/ncu_key.py <report.ncu-rep> [kernel substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    d = dict(zip(h, r))
    if sub and sub not in d.get("Kernel Name", ""): continue
    print("==", d.get("Kernel Name", "")[:80])
    for k in KEYS:
        if k in d: print("  %-90s %s" % (k, d[k]))
