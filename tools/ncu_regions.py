"""Samples per barrier-delimited region of one kernel in a (multi-kernel) ncu report; the barrier wait (charged by the sampler
to the first instructions after a BAR.SYNC) is shown separately.  usage: ncu_regions.py rep.ncu-rep <kernel-substring> [lib.so]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
n = lambda x: re.sub(r"\(int\)|\(bool\)|\s|rbt::|void", "", x)
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(secs):
    if n(kern) in n(rows[s][1]):
        e = secs[si + 1] if si + 1 < len(secs) else len(rows)
        hdr = rows[s + 1]; data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
        break
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ibar = hdr.index("stall_barrier")
iw = hdr.index("L1 Wavefronts Shared")
lines_of = None
if lib:
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    for sec in re.split(r"\n\s*\.section\s+\.text\.", txt):
        name = sec.split("\n", 1)[0]
        dem = subprocess.run(["c++filt", name.split(",")[0]], capture_output=True, text=True).stdout.strip().replace("true", "1").replace("false", "0")
        if n(kern) in n(dem):
            lines_of = []; cur = "?"
            for line in sec.split("\n"):
                mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
                if mm:
                    if "inlined at" not in line or True: cur = os.path.basename(mm.group(1)) + ":" + mm.group(2)
                    continue
                if re.search(r"/\*[0-9a-f]{4,}\*/\s+", line): lines_of.append(cur)
            break
I = lambda r, i: int(float(r[i] or 0))
tot = sum(I(r, isamp) for r in data)
print(f"total samples {tot}; instructions {len(data)}")
start = 0; regs = []
for k, r in enumerate(data):
    if "BAR.SYNC" in r[isrc] or "BAR.ARV" in r[isrc] or "BAR.RED" in r[isrc]:
        regs.append((start, k)); start = k + 1
regs.append((start, len(data) - 1))
for s, e in regs:
    c = sum(I(r, isamp) for r in data[s:e + 1]); b = sum(I(r, ibar) for r in data[s:e + 1])
    w = sum(I(r, iw) for r in data[s:e + 1]); ex = sum(I(r, iex) for r in data[s:e + 1])
    nd = sum(1 for r in data[s:e + 1] if "DMMA" in r[isrc])
    if c > 0.004 * tot:
        loc = f"{lines_of[s]} .. {lines_of[min(e, len(lines_of) - 1)]}" if lines_of else ""
        print(f" inst {s:5d}-{e:5d}: samples {100 * c / tot:5.1f} % (of which barrier wait {100 * b / tot:5.1f} %)  exec {ex:10d}  smem wavefronts {w:10d}  dmma {nd:4d}  {loc}")
