"""Markdown table of the headline metrics of one `ncu --set full` report (one column per captured kernel).
Usage: python scripts/ncu_summary.py report.ncu-rep > table.md"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__pcsamp_warps_issue_stalled_wait",
           "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
           "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
           "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_sample_count"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
names = [r[ik].split("(")[0].replace("void ", "").replace("mlfb200::", "") for r in data]
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        print(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
