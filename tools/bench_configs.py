"""Device-resident full-iteration throughput for the other BASELINE.json configs (parity-test cases; the bench line is
config 3 in bench.py): config 1 iiwa14 UnconstrOCPSolver N=50 batch 256, config 3 ANYmal jumping STO N=80 batch 512 — GPU
(CUDA events) vs the CPU oracle port (OpenMP over OCPs).  Prints one JSON line per config.  Run on the GPU box."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib
import bench
from helpers import jump_sto_schedule
from robotoc_b200 import (ANYMAL, DirectMultipleShooting, Layout, RiccatiRecursion, StageDims, StageLayout,
                          UnconstrDirectMultipleShooting, UnconstrRiccatiRecursion, anymal_constraint_table, iiwa14_constraint_table)
from robotoc_b200.layout import ULayout
from synth import make_stage_inputs, symmetrize_lin
from synth import make_unconstr_stage_inputs


def gpu_time(fn, reset, iters=10):
    for _ in range(3):
        reset(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def config_unconstr(N=50, batch=256, dt=0.02):
    tab = iiwa14_constraint_table()
    rr = UnconstrRiccatiRecursion(7, N, dt, batch)
    dms = UnconstrDirectMultipleShooting(rr, tab)
    S = dms.layout
    lin, con, sol, dx0 = make_unconstr_stage_inputs(S, N, batch, 11)
    dms.condense(lin, con); dms.setSolution(sol); rr.backwardRiccatiRecursion(); rr.forwardRiccatiRecursion(dx0)

    def reset():
        dms._up(7, con, S.c_stride, None); dms.setSolution(sol)

    def it():
        dms.condense(); rr.backwardRiccatiRecursion(); rr.forwardRiccatiRecursion(); dms.computeStepSizes(); dms.integrateSolution()
    ms = gpu_time(it, reset)
    UL = ULayout(7, getter=oracle_lib.load().orc_ulayout_get)
    oracle_lib.unconstr_iteration(7, UL, S, tab, N, dt, lin, con, sol, dx0)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 3.0:
        oracle_lib.unconstr_iteration(7, UL, S, tab, N, dt, lin, con, sol, dx0); n += 1
    cpu = batch * n / (time.perf_counter() - t0)
    rr.close()
    return {"config": f"iiwa14 UnconstrOCPSolver N={N} batch={batch} (full iteration)", "gpu_ms": ms, "gpu_ocp_iter_per_s": batch / ms * 1e3,
            "cpu_port_ocp_iter_per_s": cpu, "cpu_threads": os.cpu_count(), "note": "CPU leg includes numpy copies of the inputs"}


def config_jump_sto(N=80, batch=512):
    td, ev, ctrl = jump_sto_schedule(N)
    table = anymal_constraint_table()
    sd = StageDims(ANYMAL, nf_max=12, n_contacts=table.n_contacts, n_box=table.n_box)
    S, K = StageLayout(sd), Layout(ANYMAL)
    lin, con, sol, dx0 = make_stage_inputs(sd, S, ctrl, batch, 12)
    lin = symmetrize_lin(S, lin)
    rr = RiccatiRecursion(ANYMAL, len(ctrl), batch); rr.setTimeDiscretization(ctrl)
    dms = DirectMultipleShooting(rr, sd, table)
    dms.condense(lin, con); rr.backwardRiccatiRecursion(); rr.forwardRiccatiRecursion(dx0)

    def reset():
        dms._up(7, con, S.c_stride, None); dms._up(9, sol, S.s_stride, None)

    def it():
        dms.condense(); rr.backwardRiccatiRecursion(); rr.forwardRiccatiRecursion(); dms.computeStepSizes(); dms.integrateSolution()
    ms = gpu_time(it, reset)
    assert int(rr.info().max()) == 0
    pr = dict(dims=ANYMAL, sd=sd, S=S, K=K, table=table, ctrl=ctrl, lin=lin, con=con, sol=sol, dx0=dx0)
    cores, run = bench._oracle_runner(pr)
    el, n = 0.0, 0
    while el < 5.0:
        el += run(); n += 1
    rr.close()
    return {"config": f"ANYmal jumping STO OCPSolver N={N} ({len(ctrl)} grid points) batch={batch} (full iteration)", "gpu_ms": ms,
            "gpu_ocp_iter_per_s": batch / ms * 1e3, "cpu_port_ocp_iter_per_s": batch * n / el, "cpu_threads": cores}


if __name__ == "__main__":
    for f in (config_unconstr, config_jump_sto):
        print(json.dumps(f()), flush=True)
