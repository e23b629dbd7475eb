"""Where does a kernel spill?  Counts LDL/STL SASS instructions per source line (needs -lineinfo).
usage: python tools/spill_lines.py <mangled-kernel-name-substring> [lib.so]"""
import os, re, subprocess, sys, tempfile
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "robotoc_b200", "librobotoc_b200.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
secs = re.split(r"\n\s*\.section\s+\.text\.", txt)
for sec in secs:
    name = sec.split("\n", 1)[0]
    if pat not in name:
        continue
    cur, cnt, tot = None, {}, 0
    for line in sec.split("\n"):
        mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        if re.search(r"\b(LDL|STL)(\.\w+)*\b", line):
            cnt[cur] = cnt.get(cur, 0) + 1
            tot += 1
    print(name.split(",")[0], "LDL/STL total", tot)
    for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:25]:
        print("  ", k, v)
