#!/usr/bin/env python
"""Developer aid: executed warp-instructions per source line from an .ncu-rep."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
passes = 301989888.0 / 32
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == 'Line No': hdr = r; ia = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) <= ia or r[0] == '': continue
    try: key = (cur, int(r[0])); c = int(r[ia])
    except ValueError: continue
    agg[key] = (c, r[1].strip()[:100])
tot = sum(v[0] for v in agg.values())
print("total (with inlined double counting)", tot / passes)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:5d} {v[0] / passes:7.2f} | {v[1]}")
