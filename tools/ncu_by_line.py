"""Attribute per-instruction ncu metrics (source page, SASS) of one kernel to CUDA source lines, using the -lineinfo of the
built library (nvdisasm) -- instruction order is the same in both listings.
usage: python tools/ncu_by_line.py rep.ncu-rep <kernel-substring> <metric column, e.g. 'L1 Wavefronts Shared'> [--file f.cuh] [--top 40]"""
import argparse, collections, csv, io, os, re, subprocess, tempfile
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel"); ap.add_argument("metric"); ap.add_argument("--file", default=None)
ap.add_argument("--top", type=int, default=40); ap.add_argument("--extra", default="L1 Wavefronts Shared Excessive")
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "robotoc_b200", "librobotoc_b200.so"))
a = ap.parse_args()
src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# sections: a row ["Kernel Name", name] then header row then data
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
data = None
for si, s in enumerate(secs):
    _n = lambda x: re.sub(r"\(int\)|\(bool\)|\s|rbt::|void", "", x)
    if _n(a.kernel) in _n(rows[s][1]):
        e = secs[si + 1] if si + 1 < len(secs) else len(rows)
        hdr = rows[s + 1]
        data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
        break
assert data is not None, "kernel not found in report"
im, ix, isrc, isamp = hdr.index(a.metric), hdr.index(a.extra), hdr.index("Source"), hdr.index("# Samples")
# line map from the library
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", a.lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
mangled = None
lines_of = []
for sec in re.split(r"\n\s*\.section\s+\.text\.", txt):
    name = sec.split("\n", 1)[0]
    norm = lambda x: re.sub(r"\(int\)|\(bool\)|\s|rbt::|void", "", x)
    dem = subprocess.run(["c++filt", name.split(",")[0]], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("true", "1").replace("false", "0")
    if norm(a.kernel) in norm(dem):
        cur = ("?", 0); own = ("?", 0)
        for line in sec.split("\n"):
            mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                if a.file is None or cur[0] == a.file:
                    own = cur
                continue
            if re.search(r"/\*[0-9a-f]{4,}\*/\s+", line):
                lines_of.append(own)
        break
print(f"instructions: report {len(data)}, library {len(lines_of)}")
n = min(len(data), len(lines_of))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for k in range(n):
    try:
        v = int(float(data[k][im] or 0)); x = int(float(data[k][ix] or 0)); sm = int(data[k][isamp] or 0)
    except ValueError:
        continue
    agg[lines_of[k]][0] += v; agg[lines_of[k]][1] += x; agg[lines_of[k]][2] += sm
    tot += v
print(f"total {a.metric}: {tot}")
srcs = {}
for (f, ln), (v, x, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
    if f not in srcs:
        p = os.path.join(os.path.dirname(a.lib), "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100 * v / max(tot, 1):5.1f}%  {v:10d}  extra {100 * x / max(v, 1):5.1f}%  samples {sm:6d}  {f}:{ln:<5d} {text}")
