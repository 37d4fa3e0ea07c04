"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: python tools/ncu_summary.py report.ncu-rep [regex ...]"""
import csv, subprocess, sys, re, io
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = [r"^Kernel Name$", r"gpu__time_duration.sum$", r"launch__registers_per_thread$", r"launch__grid_size$", r"launch__block_size$",
        r"dram__bytes_read.sum$", r"dram__bytes_write.sum$", r"lts__t_bytes.sum$",
        r"sm__throughput.avg.pct_of_peak_sustained_elapsed$", r"sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active$",
        r"sm__pipe_fmaheavy_cycles_active.avg.pct", r"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active$",
        r"sm__inst_executed_pipe_xu.avg.pct", r"sm__pipe_xu_cycles_active.avg.pct",
        r"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active$",
        r"smsp__issue_active.avg.pct_of_peak_sustained_active$", r"smsp__inst_executed.sum$",
        r"smsp__warps_eligible.avg.per_cycle_active$", r"sm__warps_active.avg.pct_of_peak_sustained_active$",
        r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed$",
        r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$",
        r"sass__inst_executed_local_", r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$",
        r"sm__cycles_active.avg$", r"smsp__cycles_active.avg$", r"sm__inst_executed_pipe_[a-z_0-9]*.sum$",
        r"sm__sass_thread_inst_executed_op_ffma_pred_on.sum$", r"smsp__sass_thread_inst_executed_op_f", r"gpc__cycles_elapsed.max$",
        r"sm__cycles_elapsed.avg.per_second$"] + extra
for row in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, row):
        if any(re.search(k, h) for k in keys):
            if "stalled" in h and float(v or 0) < 0.05:
                continue
            print(f"{h:100s} {v} {u}")
