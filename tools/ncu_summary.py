"""Summarise an .ncu-rep (raw + source pages) into a short text: key metrics + stall samples per barrier-delimited region."""
import csv, subprocess, sys, bisect, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg"]
print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:80s} {v:>18s} {u}")
st = {h: v for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")}
tot = sum(float(v) for v in st.values()) or 1
print("== stall reasons (pc samples):")
for h, v in sorted(st.items(), key=lambda kv: -float(kv[1]))[:8]:
    print(f"   {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {float(v) / tot * 100:5.1f} %")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data) or 1
regions = []; cur = 0; start = 0; loc = 0
for n, r in enumerate(data):
    cur += int(r[isamp])
    if "LDL" in r[isrc] or "STL" in r[isrc]:
        loc += int(r[iex])
    if "BAR.SYNC" in r[isrc] or "BAR.ARV" in r[isrc]:
        regions.append((start, n, cur, loc)); cur = 0; loc = 0; start = n + 1
regions.append((start, len(data) - 1, cur, loc))
print(f"== samples between barriers (total {tot}); a barrier wait is charged to the first instruction after it")
for s, e, c, l in regions:
    if c > tot * 0.01:
        top = max(data[s:e + 1], key=lambda r: int(r[isamp]))
        nd = sum(1 for r in data[s:e + 1] if "DMMA" in r[isrc])
        print(f"   inst {s:6d}-{e:6d}  {100 * c / tot:5.1f} %  dmma={nd:4d} local_exec={l:9d}  top: {top[isrc].strip()[:48]} ({top[isamp]})")
