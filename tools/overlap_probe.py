"""Does co-scheduling help?  Full iteration for 1024 OCPs as (a) one handle, one stream; (b) two 512-OCP handles on one
stream (sequential); (c) the two handles on two streams, B delayed so that condense(B) overlaps backward(A)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import trot_schedule
from robotoc_b200 import ANYMAL, DirectMultipleShooting, RiccatiRecursion, StageDims, StageLayout, anymal_constraint_table
from synth import make_stage_inputs

td, ev, ctrl = trot_schedule(40)
table = anymal_constraint_table()
sd = StageDims(ANYMAL, nf_max=12, n_contacts=table.n_contacts, n_box=table.n_box)
S = StageLayout(sd)


def make(batch, seed):
    lin, con, sol, dx0 = make_stage_inputs(sd, S, ctrl, batch, seed)
    rr = RiccatiRecursion(ANYMAL, len(ctrl), batch); rr.setTimeDiscretization(ctrl)
    dms = DirectMultipleShooting(rr, sd, table)
    dms.condense(lin, con); dms._up(9, sol, S.s_stride, None); rr.backwardRiccatiRecursion(); rr.forwardRiccatiRecursion(dx0); rr.synchronize()
    return rr, dms, con, sol


def iteration(rr, dms, sp):
    dms.condense(stream=sp); rr.backwardRiccatiRecursion(stream=sp); rr.forwardRiccatiRecursion(stream=sp)
    dms.computeStepSizes(stream=sp); dms.integrateSolution(stream=sp)


def timeit(fn, n=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

full = make(1024, 1)
A, B = make(512, 2), make(512, 3)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
p1, p2 = ctypes.c_void_p(s1.cuda_stream), ctypes.c_void_p(s2.cuda_stream)
cur = torch.cuda.current_stream()


def one():
    iteration(full[0], full[1], None)


def seq():
    iteration(A[0], A[1], None); iteration(B[0], B[1], None)


def par():
    s1.wait_stream(cur); s2.wait_stream(cur)
    iteration(A[0], A[1], p1); iteration(B[0], B[1], p2)
    cur.wait_stream(s1); cur.wait_stream(s2)


def staggered():
    # A: condense | backward ...   B starts its condense when A's condense is done -> condense(B) || backward(A)
    s1.wait_stream(cur); s2.wait_stream(cur)
    A[1].condense(stream=p1)
    ev_ = torch.cuda.Event(); ev_.record(s1); s2.wait_event(ev_)
    A[0].backwardRiccatiRecursion(stream=p1); B[1].condense(stream=p2)
    A[0].forwardRiccatiRecursion(stream=p1); A[1].computeStepSizes(stream=p1); A[1].integrateSolution(stream=p1)
    B[0].backwardRiccatiRecursion(stream=p2); B[0].forwardRiccatiRecursion(stream=p2); B[1].computeStepSizes(stream=p2); B[1].integrateSolution(stream=p2)
    cur.wait_stream(s1); cur.wait_stream(s2)

for name, fn in (("one handle 1024", one), ("two x 512 sequential", seq), ("two x 512 on two streams", par), ("two x 512 staggered", staggered)):
    print(f"{name:28s} {timeit(fn):.3f} ms", flush=True)
