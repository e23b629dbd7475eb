import csv, subprocess, sys, io
rep, kidx = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + kidx]
print("==", vals[hdr.index("Kernel Name")])
keys = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","launch__waves_per_multiprocessor",
 "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__cycles_active.avg",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__sass_inst_executed_op_local_ld.sum","smsp__sass_inst_executed_op_local_st.sum",
 "smsp__inst_executed_pipe_lsu.sum","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
 "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct","l1tex__throughput.avg.pct_of_peak_sustained_active"]
for h,u,v in zip(hdr,units,vals):
    if h in keys: print(f"{h:75s} {v:>16s} {u}")
st = {h: v for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")}
tot = sum(float(v) for v in st.values()) or 1
for h, v in sorted(st.items(), key=lambda kv: -float(kv[1]))[:10]:
    print(f"   {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {float(v) / tot * 100:5.1f} %")
