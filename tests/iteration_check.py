"""TEST INFRASTRUCTURE: one full hot-path iteration on the CPU oracle and the block-by-block comparison of a CUDA result
with it.  Shared by the -m gpu parity tests, __graft_entry__.smoke() and the pre-timing check of bench.py.

  condense -> backward Riccati -> forward Riccati -> expand + step sizes -> update
  (the linear-algebra body of OCPSolver::updateSolution, /root/reference/src/solver/ocp_solver.cpp:118-144)
"""
import ctypes

import numpy as np

import oracle_lib
from robotoc_b200 import ANYMAL
from robotoc_b200.grid import IMPACT, TERMINAL


def oracle_iteration(sd, S, K, table, ctrl, lin, con, sol, dx0, nthreads=0):
    """Returns every intermediate record of the iteration as the oracle computes it (inputs are not modified)."""
    lib = oracle_lib.load()
    P = oracle_lib.ptr
    batch, n_grid = lin.shape[0], lin.shape[1]
    csd = sd.c()
    kkt = np.zeros((batch, n_grid, K.k_stride))
    ex = np.zeros((batch, n_grid, S.e_stride))
    cc, ss = con.copy(), sol.copy()
    rc = lib.orc_condense_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, P(lin), P(cc), P(kkt), P(ex), nthreads)
    assert rc == 0, "oracle: condensing failed (non-SPD M / J M^-1 J^T)"
    cc_cond = cc.copy()
    kk, ric, d, info = oracle_lib.riccati_batch(ANYMAL, K, ctrl, kkt, dx0, nthreads=nthreads)
    assert info == 0, "oracle: Cholesky failure in the Riccati recursion"
    xd = np.zeros((batch, n_grid, S.x_stride))
    steps = np.zeros((batch, 2))
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, P(lin), P(ex), P(d), P(cc), P(xd), P(steps), nthreads)
    d_exp, cc_exp, xd_exp = d.copy(), cc.copy(), xd.copy()
    lib.orc_update_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, P(ex), P(d), P(xd), P(cc), P(ss), P(steps), nthreads)
    return dict(kkt=kkt, cc_cond=cc_cond, ric=ric, d=d_exp, cc_exp=cc_exp, xd_exp=xd_exp, steps=steps, d_upd=d, xd_upd=xd,
                cc_upd=cc, sol=ss, ex_upd=ex)


def _cmp(name, got, ref, tol, worst):
    scale = float(np.max(np.abs(ref)))
    if scale == 0.0:
        assert float(np.max(np.abs(got))) == 0.0, f"{name}: expected zeros"
        return
    e = float(np.max(np.abs(got - ref))) / scale
    worst[0] = max(worst[0], e)
    assert e < tol, f"{name}: rel err {e:.3e} (tol {tol:g})"


def compare_final(S, K, ctrl, ref, ric, d, steps, sol, cc, tol=1e-8):
    """What a caller of the iteration sees at its end, EVERY OCP of the batch, block by block and stage by stage:
    P, s, K, k (north_star: 1e-6 relative; asserted at `tol`), the Newton direction dx, du, dlmd|dgmm (after the costate
    correction of the update), the horizon-wide step sizes, the updated solution and slack / dual.  Returns the worst error."""
    nx, nu = K.nx, K.nu
    worst = [0.0]
    for i, c in enumerate(ctrl):
        _cmp(f"P[{i}]", ric[:, i, K.r_P:K.r_P + nx * nx], ref["ric"][:, i, K.r_P:K.r_P + nx * nx], tol, worst)
        _cmp(f"s[{i}]", ric[:, i, K.r_s:K.r_s + nx], ref["ric"][:, i, K.r_s:K.r_s + nx], tol, worst)
        _cmp(f"dx[{i}]", d[:, i, K.d_dx:K.d_dx + nx], ref["d_upd"][:, i, K.d_dx:K.d_dx + nx], tol, worst)
        _cmp(f"dlmdgmm[{i}]", d[:, i, K.d_dlmdgmm:K.d_dlmdgmm + nx], ref["d_upd"][:, i, K.d_dlmdgmm:K.d_dlmdgmm + nx], tol, worst)
        _cmp(f"sol[{i}]", sol[:, i], ref["sol"][:, i], tol, worst)
        if c.type in (IMPACT, TERMINAL):
            continue
        _cmp(f"K[{i}]", ric[:, i, K.r_K:K.r_K + nx * nu], ref["ric"][:, i, K.r_K:K.r_K + nx * nu], tol, worst)
        _cmp(f"k[{i}]", ric[:, i, K.r_k:K.r_k + nu], ref["ric"][:, i, K.r_k:K.r_k + nu], tol, worst)
        _cmp(f"du[{i}]", d[:, i, K.d_du:K.d_du + nu], ref["d_upd"][:, i, K.d_du:K.d_du + nu], tol, worst)
        if c.ns > 0:
            _cmp(f"dxi[{i}]", d[:, i, K.d_dxi:K.d_dxi + c.ns], ref["d_upd"][:, i, K.d_dxi:K.d_dxi + c.ns], tol, worst)
        for f in ("c_slack", "c_dual"):
            o = getattr(S, f)
            _cmp(f"{f}[{i}]", cc[:, i, o:o + S.nc], ref["cc_upd"][:, i, o:o + S.nc], tol, worst)
    _cmp("steps", steps, ref["steps"], 1e-10, worst)
    return worst[0]


def run_device_iteration(rr, dms, lin, con, sol, dx0, stream=None):
    """The five C-ABI calls of one iteration on device-resident records; returns what compare_final needs."""
    dms.condense(lin, con, stream=stream)
    rr.backwardRiccatiRecursion(stream=stream)
    rr.forwardRiccatiRecursion(dx0, stream=stream)
    dms.computeStepSizes(stream=stream)
    dms.integrateSolution(sol, stream=stream)
    steps = np.stack([dms.maxPrimalStepSize(stream), dms.maxDualStepSize(stream)], axis=1)
    return dict(ric=rr.getRiccatiFactorization(stream), d=rr.getDirection(stream), steps=steps, sol=dms.getSolution(stream),
                cc=dms.getConstraintData(stream), info=rr.info(stream))
