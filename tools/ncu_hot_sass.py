#!/usr/bin/env python
"""Developer aid: the executed hot path of a kernel from an .ncu-rep, instruction by instruction --
SASS with executed warp-instructions and shared-memory wavefronts per record-chunk (one wave
operator x one 128-sample chunk) and warp-stall samples; lines executed less than `min_share`
times per record-chunk are left out.  Written to profiles/ as the SASS excerpt of the chunk loop.
usage: tools/ncu_hot_sass.py gpurun_out/X.ncu-rep out.txt [record_chunks_per_launch] [min_share]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
rc = float(sys.argv[3]) if len(sys.argv) > 3 else 3 * 4096 * 24576 / 128.0
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Instructions Executed" in r and "Source" in r)
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
iw, ist = hdr.index("L1 Wavefronts Shared"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[rows.index(hdr) + 1:]:
    try:
        data.append((r[isrc].strip(), int(r[ia]), int(r[iw] or 0), int(r[ist] or 0)))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data)
hot = [(i, d) for i, d in enumerate(data) if d[1] >= min_share * rc]
byop = collections.Counter()
for _, d in hot:
    t = d[0].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += d[1]
with open(out, "w") as f:
    f.write(f"# {rep}: executed hot path (>= {min_share} executions per record-chunk; record-chunk = one wave "
            f"operator x one 128-sample chunk, {rc:.0f} per launch)\n")
    f.write(f"# all instructions: {tot / rc:.1f} per record-chunk = {tot / rc / 4:.1f} per 32 op-samples; "
            f"listed: {sum(d[1] for _, d in hot) / rc:.1f}; shared-memory wavefronts listed: "
            f"{sum(d[2] for _, d in hot) / rc:.1f} per record-chunk\n")
    f.write("# by opcode (per record-chunk): " + ", ".join(f"{k} {v / rc:.1f}" for k, v in byop.most_common(24)) + "\n")
    f.write("# index  executions/record-chunk  smem wavefronts/record-chunk  stall samples  SASS\n")
    prev = None
    for i, d in hot:
        if prev is not None and i != prev + 1:
            f.write("   ....\n")
        f.write(f"{i:6d} {d[1] / rc:6.2f} w{d[2] / rc:6.2f} st{d[3]:5d}  {d[0]}\n")
        prev = i
print("wrote", out, len(hot), "instructions")
