"""Step-by-step GPU bring-up (each step prints and flushes so a hang is attributable)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib
from helpers import small_event_schedule, trot_schedule, rel_err
from robotoc_b200 import ANYMAL, Layout, RiccatiRecursion, ULayout, UnconstrRiccatiRecursion
from robotoc_b200.schedule import plain_schedule
from robotoc_b200.synth import make_kkt, make_unconstr_kkt

step = sys.argv[1]
def log(*a):
    print(*a, flush=True)

if step == "unconstr":
    nv, N, dt, batch = 7, 20, 0.05, 2
    UL = ULayout(nv)
    kkt, dx0 = make_unconstr_kkt(nv, UL, N, batch, 1)
    ur = UnconstrRiccatiRecursion(nv, N, dt, batch)
    log("created"); ur.backwardRiccatiRecursion(kkt); ur.synchronize(); log("bwd done")
    ur.forwardRiccatiRecursion(dx0); ur.synchronize(); log("fwd done")
    ric, d = ur.getRiccatiFactorization(), ur.getDirection()
    kk, ro, do, info = oracle_lib.unconstr_batch(nv, UL, N, dt, kkt, dx0)
    log("P err", rel_err(ric[..., :196], ro[..., :196]), "dir err", rel_err(d, do))
else:
    dims = ANYMAL; L = Layout(dims)
    if step == "plain":
        ctrl = plain_schedule(6, 0.03, 12)
    elif step == "event":
        td, ev, ctrl = small_event_schedule(False)
    elif step == "trot":
        td, ev, ctrl = trot_schedule(40)
    batch = 3
    kkt, dx0 = make_kkt(dims, L, ctrl, batch, 1)
    rr = RiccatiRecursion(dims, len(ctrl), batch); rr.setTimeDiscretization(ctrl)
    log("created")
    rr.backwardRiccatiRecursion(kkt, write_fact=True); rr.synchronize(); log("bwd done")
    ric = rr.getRiccatiFactorization()
    kk, ro, do, info = oracle_lib.riccati_batch(dims, L, ctrl, kkt, dx0)
    nx, nu = dims.nx, dims.nu
    for i in range(len(ctrl) - 1, -1, -1):
        log(i, ctrl[i].type, ctrl[i].ns, "P", "%.2e" % rel_err(ric[:, i, L.r_P:L.r_P + nx * nx], ro[:, i, L.r_P:L.r_P + nx * nx]),
            "s", "%.2e" % rel_err(ric[:, i, L.r_s:L.r_s + nx], ro[:, i, L.r_s:L.r_s + nx]),
            "K", "%.2e" % rel_err(ric[:, i, L.r_K:L.r_K + nx * nu], ro[:, i, L.r_K:L.r_K + nx * nu] + 1e-300),
            "k", "%.2e" % rel_err(ric[:, i, L.r_k:L.r_k + nu], ro[:, i, L.r_k:L.r_k + nu] + 1e-300))
    rr.forwardRiccatiRecursion(dx0); rr.synchronize(); log("fwd done")
    d = rr.getDirection()
    for i in range(len(ctrl)):
        log(i, "dx %.2e" % rel_err(d[:, i, L.d_dx:L.d_dx + nx], do[:, i, L.d_dx:L.d_dx + nx]),
            "du %.2e" % rel_err(d[:, i, L.d_du:L.d_du + nu], do[:, i, L.d_du:L.d_du + nu] + 1e-300),
            "dl %.2e" % rel_err(d[:, i, L.d_dlmdgmm:L.d_dlmdgmm + nx], do[:, i, L.d_dlmdgmm:L.d_dlmdgmm + nx]))
    log("info", rr.info())
