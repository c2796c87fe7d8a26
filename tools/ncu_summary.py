#!/usr/bin/env python
"""Developer aid: summarise an .ncu-rep of render_kernel here (no GPU needed).
usage: tools/ncu_summary.py gpurun_out/X.ncu-rep [op_samples_per_launch] [nlines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
opsamp = float(sys.argv[2]) if len(sys.argv) > 2 else 301989888.0
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 40
passes = opsamp / 32.0

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k]}")
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(m[k])
        if v > 0.15:
            print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
tot = float(m["smsp__inst_executed.sum"])
print(f"warp-instructions per 32 op-samples: {tot / passes:.1f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
byop = collections.Counter()
nstatic = 0
for r in rows:
    if "Instructions Executed" in r and "Source" in r:
        h = r
        ia, isrc = h.index("Instructions Executed"), h.index("Source")
        continue
    if h is None or len(r) <= ia:
        continue
    try:
        cnt = int(r[ia])
    except ValueError:
        continue
    t = r[isrc].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    parts = op.split(".")
    op = ".".join(parts[:3]) if parts[0] in ("F2F", "I2F", "F2I", "MUFU", "LDS", "STS", "LDL", "STL", "LDG", "STG", "I2FP") else parts[0]
    byop[op] += cnt
    nstatic += 1
print("static SASS instructions:", nstatic)
for op, cnt in byop.most_common(nlines):
    print(f"  {op:18s} {cnt / passes:8.2f} per op-sample pass  {100 * cnt / tot:5.1f}%")
