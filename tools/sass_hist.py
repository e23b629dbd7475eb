"""Static SASS instruction count per source region of a kernel (needs -lineinfo): instructions from inlined helpers are
charged to the last line of `--file` seen before them.  usage: python tools/sass_hist.py <kernel-substring> --file riccati_backward.cuh [--bucket 10]"""
import argparse, collections, os, re, subprocess, tempfile
ap = argparse.ArgumentParser(); ap.add_argument("kernel"); ap.add_argument("--file", required=True); ap.add_argument("--bucket", type=int, default=1)
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "robotoc_b200", "librobotoc_b200.so"))
a = ap.parse_args()
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", a.lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
for sec in re.split(r"\n\s*\.section\s+\.text\.", txt):
    if a.kernel not in sec.split("\n", 1)[0]:
        continue
    cur, hist, ops = 0, collections.Counter(), collections.defaultdict(collections.Counter)
    for line in sec.split("\n"):
        mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if mm:
            if os.path.basename(mm.group(1)) == a.file:
                cur = int(mm.group(2)) // a.bucket * a.bucket
            continue
        m2 = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m2:
            hist[cur] += 1
            ops[cur][m2.group(2).split(".")[0]] += 1
    print(sec.split("\n", 1)[0].split(",")[0], "total", sum(hist.values()))
    for k in sorted(hist):
        if hist[k] >= 20:
            print(f"  line {k:5d}: {hist[k]:5d}  " + " ".join(f"{o}:{n}" for o, n in ops[k].most_common(6)))
