"""Shared-memory wavefronts per CUDA source line (total / excessive = bank-conflict replays).  usage: ncu_smem.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
recs = []; hdr = None
for r in rows:
    if "# Samples" in r:
        hdr = r; iw = hdr.index("L1 Wavefronts Shared"); ie = hdr.index("L1 Wavefronts Shared Excessive"); continue
    if hdr and len(r) > iw and r[0].strip().isdigit():
        try:
            w = int(r[iw] or 0); e = int(r[ie] or 0)
        except ValueError:
            continue
        if w:
            recs.append((w, e, r[0], r[1].strip()))
tw = sum(x[0] for x in recs) or 1; te = sum(x[1] for x in recs)
print(f"total shared wavefronts {tw}, excessive {te} ({100 * te / tw:.1f} %)")
for w, e, ln, s in sorted(recs, key=lambda x: -x[0])[:topn]:
    print(f"{100 * w / tw:5.1f}%  excess {100 * e / max(w, 1):5.1f}%  :{ln:>5s}  {s[:110]}")
