"""Per-phase cycle budget of one CTA of the backward kernel (RBT_TIMELINE_CTA instrumentation)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("RBT_TIMELINE_CTA", "300")
from bench import build_problem
from robotoc_b200 import RiccatiRecursion
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dims, L, ctrl, kkt, dx0 = build_problem(batch, 1)
rr = RiccatiRecursion(dims, len(ctrl), batch); rr.setTimeDiscretization(ctrl)
rr.backwardRiccatiRecursion(kkt); rr.synchronize()
rr.backwardRiccatiRecursion(); rr.synchronize()
n = len(ctrl) * 32
buf = (ctypes.c_longlong * n)()
rr._lib.rbt_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
assert rr._lib.rbt_debug_timeline(rr._h, buf, n) == 0
tl = np.array(buf[:], dtype=np.int64).reshape(len(ctrl), 2, 16)
names_g = ["wait0", "tma_done", "gemm1", "bar1", "gemm2", "H", "t1", "arr_b2", "b2", "phaseC", "b3", "phaseD+spill", "b4", "phaseE", "b5"]
for st in (40, 30, 19, 10):
    g, f = tl[st, 0], tl[st, 1]
    print(f"stage {st} type {ctrl[st].type} ns {ctrl[st].ns}: GEMM warp0 deltas:", " ".join(f"{names_g[k]}={g[k]-g[k-1]}" for k in range(1, 15) if g[k] and g[k-1]))
    print(f"          factor warp: tma_done={f[1]-f[0]} Bp={f[2]-f[1]} G={f[3]-f[2]} chol={f[4]-f[3]} wait_z={f[5]-f[4]} lu'={f[7]-f[5]} b2={f[8]-f[7]} C(Linv)={f[9]-f[8]} b3={f[10]-f[9]} D(K)={f[11]-f[10]} b4={f[12]-f[11]} E={f[13]-f[12]} b5={f[14]-f[13]}")
    print(f"          stage total (GEMM warp0 b5->b5): {tl[st,0,14]-tl[st+1,0,14]} cycles")
tot = tl[0, 0, 14] - tl[len(ctrl) - 2, 0, 0]
print("sweep total cycles", tot, "=> per stage", tot / (len(ctrl) - 1))
