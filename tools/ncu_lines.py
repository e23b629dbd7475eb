"""Aggregate pc samples of an .ncu-rep by CUDA source line (needs -lineinfo + --import-source on).  usage: ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname = ""
recs = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if "# Samples" in r:
        hdr = r; isamp = hdr.index("# Samples"); continue
    if hdr and len(r) > isamp and r[0].strip().isdigit() and r[isamp].isdigit() and int(r[isamp]) > 0:
        recs.append((int(r[isamp]), fname, r[0], r[1].strip()))
tot = sum(x[0] for x in recs) or 1
print("total samples", tot)
for n, f, ln, s in sorted(recs, key=lambda x: -x[0])[:topn]:
    print(f"{100 * n / tot:5.1f}%  {f}:{ln:>5s}  {s[:120]}")
