"""Brief summary of an .ncu-rep: duration, pipes, issue, stall mix, memory.  usage: python tools/ncu_brief.py report.ncu-rep"""
import csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [r"^gpu__time_duration.sum$", r"^launch__registers_per_thread$", r"^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active$",
        r"^sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active$",
        r"^sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active$",
        r"^smsp__issue_active.avg.pct_of_peak_sustained_active$", r"^smsp__inst_executed.sum$", r"^dram__bytes_read.sum$", r"^dram__bytes_write.sum$",
        r"^lts__t_bytes.sum$", r"^lts__throughput.avg.pct_of_peak_sustained_elapsed$", r"^lts__t_sectors_srcunit_tex_op_read.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"^sm__cycles_elapsed.avg.per_second$", r"^sm__cycles_active.avg$",
        r"^smsp__pcsamp_warps_issue_stalled_[a-z_]*(?<!not_issued)$", r"^sm__throughput.avg.pct_of_peak_sustained_elapsed$"]
for row in rows[2:]:
    print("=" * 80)
    stalls = []
    for h, u, v in zip(hdr, units, row):
        if any(re.search(k, h) for k in want):
            if "pcsamp" in h:
                stalls.append((float(v or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            else:
                print(f"{h:85s} {v} {u}")
    tot = sum(s for s, _ in stalls) or 1
    print("stall samples:", ", ".join(f"{n} {100*s/tot:.1f}%" for s, n in sorted(stalls, reverse=True)[:9]))
