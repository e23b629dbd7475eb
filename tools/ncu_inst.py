"""Executed warp-instructions per CUDA source line.  usage: ncu_inst.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
recs = []; hdr = None
for r in rows:
    if "# Samples" in r:
        hdr = r; ii = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ii and r[0].strip().isdigit():
        try:
            n = int(r[ii] or 0)
        except ValueError:
            continue
        if n:
            recs.append((n, r[0], r[1].strip()))
tot = sum(x[0] for x in recs) or 1
print("total executed warp-instructions", tot)
for n, ln, s in sorted(recs, key=lambda x: -x[0])[:topn]:
    print(f"{100 * n / tot:5.1f}%  :{ln:>5s}  {s[:120]}")
