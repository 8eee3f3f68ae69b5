#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one ncu --set full capture (needs -lineinfo):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ix = hdr.index("Instructions Executed"); sx = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        n = int(r[ix]); s = int(r[sx])
    except ValueError:
        continue
    agg[(cur_file, int(r[0]), r[1].strip()[:110])] = (n, s)
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot}, samples {ts}")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/max(ts,1):5.1f}% stall  {f}:{ln}  {src}")
