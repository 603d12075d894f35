"""Summarise an ncu report: headline metrics + instructions / stall samples per source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__shared_mem_per_block_dynamic"]
for i, n in enumerate(h):
    if n in want or ("issue_stalled" in n and "per_issue_active" in n and float(v[i] or 0) > 0.2):
        print(f"{n:85s} {rows[1][i]:>12s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hdr, out = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
        ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r and r[0].isdigit() and len(r) > ie:
        try:
            out.append((cur.split("/")[-1], int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[si] or 0)))
        except ValueError:
            pass
ti, ts = sum(o[3] for o in out), sum(o[4] for o in out)
print("total warp-inst", ti, "samples", ts)
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{o[0][:12]:12s}{o[1]:5d} inst {o[3]/ti*100:5.1f}% samp {o[4]/ts*100:5.1f}%  {o[2][:100]}")
