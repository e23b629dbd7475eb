"""Pins oracle/condense_oracle.c (rows a10-a16) by identities that do not reuse its formulas:
 * MJtJinv == dense inverse of [[M, J^T],[J, 0]]                         (robot.hxx:642-683)
 * expansion satisfies the linearised contact dynamics it eliminated     (contact_dynamics.cpp:167-174)
 * the condensed quadratic model == the uncondensed model with (a, f) substituted   (contact_dynamics.cpp:55-128)
 * PDIPM: box / friction-cone condensing == J^T diag(z/s) J and J^T cond   (joint_*_limit.cpp, friction_cone.cpp:194-235)
 * fraction-to-boundary keeps slack, dual positive; update == x + alpha dx
"""
import ctypes

import numpy as np
import pytest

import oracle_lib
from helpers import small_event_schedule
from robotoc_b200 import ANYMAL, Layout
from robotoc_b200.grid import IMPACT, TERMINAL, plain_schedule
from robotoc_b200.stage import StageDims, StageLayout, anymal_constraint_table, rbt_constraint_table
from synth import make_stage_inputs
from synth import mat


def _setup(table, ctrl, batch=2, seed=5):
    lib = oracle_lib.load()
    sd = StageDims(ANYMAL, nf_max=12, n_contacts=table.n_contacts, n_box=table.n_box)
    S = StageLayout(sd, getter=lib.orc_stage_layout_get)
    K = Layout(ANYMAL, getter=lib.orc_layout_get)
    lin, con, sol, dx0 = make_stage_inputs(sd, S, ctrl, batch, seed)
    return lib, sd, S, K, lin, con, sol, dx0


def _condense(lib, sd, S, K, table, ctrl, lin, con):
    batch, n_grid = lin.shape[0], lin.shape[1]
    kkt = np.zeros((batch, n_grid, K.k_stride))
    ex = np.zeros((batch, n_grid, S.e_stride))
    cc = con.copy()
    csd = sd.c()
    info = lib.orc_condense_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(lin),
                                  oracle_lib.ptr(cc), oracle_lib.ptr(kkt), oracle_lib.ptr(ex), 1)
    assert info == 0
    return kkt, ex, cc


def test_mjtjinv_is_dense_inverse():
    lib = oracle_lib.load()
    rng = np.random.default_rng(0)
    nv, nf = 18, 12
    Sm = rng.uniform(-1, 1, (nv, nv))
    M = np.eye(nv) + 0.1 * Sm @ Sm.T
    J = rng.uniform(-1, 1, (nf, nv))
    Z = np.zeros((nv + nf) * (nv + nf))
    Mc, Jc = np.asfortranarray(M).ravel(order="F").copy(), np.asfortranarray(J).ravel(order="F").copy()
    assert lib.orc_mjtjinv(nv, nf, oracle_lib.ptr(Mc), oracle_lib.ptr(Jc), nf, oracle_lib.ptr(Z), nv + nf) == 0
    full = np.block([[M, J.T], [J, np.zeros((nf, nf))]])
    assert np.allclose(Z.reshape(nv + nf, nv + nf, order="F"), np.linalg.inv(full), rtol=1e-10, atol=1e-12)


def test_condensed_model_equals_substituted_model():
    table = rbt_constraint_table()  # no inequality rows: isolates the dynamics condensing
    table.barrier, table.fraction_to_boundary = 1e-3, 0.995
    ctrl = plain_schedule(3, 0.03, 12)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl)
    kkt, ex, _ = _condense(lib, sd, S, K, table, ctrl, lin, con)
    nv, nu, nx, npass, nf, nvfm = 18, 12, 36, 6, 12, 30
    rng = np.random.default_rng(1)
    for b in range(2):
        l, k, e = lin[b, 1], kkt[b, 1], ex[b, 1]
        M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, nf, nv)
        D, IDC = mat(l, S.l_D, nvfm, nx), l[S.l_IDC:S.l_IDC + nvfm]
        Qxx0, Quu0 = mat(l, S.l_Qxx, nx, nx), mat(l, S.l_Quu, nu, nu)
        Qaa, Qff, Qqf = np.diag(l[S.l_Qaa:S.l_Qaa + nv]), mat(l, S.l_Qff, nf, nf), mat(l, S.l_Qqf, nv, nf)
        lx0, la, lf, lu0 = (l[S.l_lx:S.l_lx + nx], l[S.l_la:S.l_la + nv], l[S.l_lf:S.l_lf + nf], l[S.l_lu:S.l_lu + nu])
        full_inv = np.linalg.inv(np.block([[M, J.T], [J, np.zeros((nf, nf))]]))
        Sel = np.zeros((nv + nf, nu))
        Sel[npass:nv, :] = np.eye(nu)

        def elim(x, u):  # (a, f) from the linearised inverse dynamics + contact constraint
            ag = full_inv @ (-D @ x + Sel @ u - IDC)
            return ag[:nv], -ag[nv:]

        def phi(x, u):
            a, f = elim(x, u)
            return (0.5 * x @ Qxx0 @ x + 0.5 * a @ Qaa @ a + 0.5 * f @ Qff @ f + x[:nv] @ Qqf @ f + 0.5 * u @ Quu0 @ u
                    + lx0 @ x + la @ a + lf @ f + lu0 @ u)

        Qxx, Qxu, Quu = mat(k, K.k_Qxx, nx, nx), mat(k, K.k_Qxu, nx, nu), mat(k, K.k_Quu, nu, nu)
        lx, lu = k[K.k_lx:K.k_lx + nx], k[K.k_lu:K.k_lu + nu]

        def psi(x, u):
            return 0.5 * x @ Qxx @ x + x @ Qxu @ u + 0.5 * u @ Quu @ u + lx @ x + lu @ u

        for _ in range(4):
            x0, u0, x1, u1 = rng.uniform(-1, 1, nx), rng.uniform(-1, 1, nu), rng.uniform(-1, 1, nx), rng.uniform(-1, 1, nu)
            lhs, rhs = phi(x1, u1) - phi(x0, u0), psi(x1, u1) - psi(x0, u0)
            assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
        # state equation rows produced by condensing: dv+ = dv + dt*da  =>  Fvq, Fvv, Fvu, Fv (contact_dynamics.cpp:130-135)
        dt = ctrl[1].dt
        x, u = rng.uniform(-1, 1, nx), rng.uniform(-1, 1, nu)
        a, _ = elim(x, u)
        Fxx, Fvu, Fx = mat(k, K.k_Fxx, nx, nx), mat(k, K.k_Fvu, nv, nu), k[K.k_Fx:K.k_Fx + nx]
        Fv0 = l[S.l_Fx + nv:S.l_Fx + nx]
        assert np.allclose((Fxx @ x)[nv:] + Fvu @ u + Fx[nv:], x[nv:] + dt * a + Fv0, rtol=1e-10, atol=1e-12)


def test_pdipm_condensing_is_JtWJ():
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(False)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=1, seed=9)
    # compare against the same stage condensed WITHOUT inequality rows: the difference must be J^T W J / J^T cond pushed
    # through the (linear) dynamics condensing.  Checked on Quu / lu, which only the torque limits touch before condensing.
    empty = rbt_constraint_table()
    empty.barrier, empty.fraction_to_boundary = table.barrier, table.fraction_to_boundary
    sd0 = StageDims(ANYMAL, nf_max=12, n_contacts=0, n_box=0)
    S0 = StageLayout(sd0, getter=lib.orc_stage_layout_get)
    i = 3  # an Intermediate stage with two feet on the ground
    assert ctrl[i].type == 0 and ctrl[i].nf == 6
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    slack, dual, res = (con[0, i, S.c_slack:S.c_slack + S.nc], con[0, i, S.c_dual:S.c_dual + S.nc],
                        con[0, i, S.c_res:S.c_res + S.nc])
    cmpl = slack * dual - table.barrier
    cond = (dual * res - cmpl) / slack
    active = np.ones(S.nc, bool)
    for ci in range(4):
        if not (ctrl[i].contact_mask >> ci) & 1:
            active[72 + 5 * ci:72 + 5 * ci + 5] = False
    assert np.allclose(cc[0, i, S.c_cmpl:S.c_cmpl + S.nc][active], cmpl[active])
    assert np.allclose(cc[0, i, S.c_cond:S.c_cond + S.nc][active], cond[active])
    assert np.all(cc[0, i, S.c_cond:S.c_cond + S.nc][~active] == 0.0)
    # torque limits rows 48..71: lower (sign -1) then upper (+1) on u[0..11]
    w = dual / slack
    dQuu = np.diag(w[48:60] + w[60:72])
    dlu = -cond[48:60] + cond[60:72]
    lin0 = np.zeros((1, len(ctrl), S0.l_stride))
    lin0[..., :min(S0.l_stride, S.l_stride)] = lin[..., :min(S0.l_stride, S.l_stride)]
    # zero-row table shares the leading sections of the linearization record (dg/dq blocks come last)
    for f in ("l_M", "l_J", "l_D", "l_IDC", "l_Qaa", "l_Qff", "l_Qqf", "l_Qxx", "l_Quu", "l_lx", "l_la", "l_lf", "l_lu", "l_Fx"):
        assert getattr(S0, f) == getattr(S, f)
    lin_mod = lin.copy()
    Quu_in = mat(lin_mod[0, i], S.l_Quu, 12, 12)
    Quu_in += dQuu
    lin_mod[0, i, S.l_lu:S.l_lu + 12] += dlu
    # strip every other row's effect by zeroing their weights: emulate with a table holding only the torque rows
    tq = rbt_constraint_table()
    tq.barrier, tq.fraction_to_boundary, tq.n_contacts = table.barrier, table.fraction_to_boundary, 0
    tq.n_box = 24
    for r in range(24):
        tq.box[r].var, tq.box[r].idx, tq.box[r].sign = 3, r % 12, (-1 if r < 12 else 1)
    sdq = StageDims(ANYMAL, nf_max=12, n_contacts=0, n_box=24)
    Sq = StageLayout(sdq, getter=lib.orc_stage_layout_get)
    linq = np.zeros((1, len(ctrl), Sq.l_stride))
    linq[..., :Sq.l_dgdq] = lin[..., :Sq.l_dgdq]
    conq = np.zeros((1, len(ctrl), Sq.c_stride))
    for name in ("c_slack", "c_dual", "c_res"):
        conq[0, :, getattr(Sq, name):getattr(Sq, name) + 24] = con[0, :, getattr(S, name) + 48:getattr(S, name) + 72]
    kq, _, _ = _condense(lib, sdq, Sq, K, tq, ctrl, linq, conq)
    lin0[..., :S0.l_dgdq] = lin_mod[..., :S0.l_dgdq]
    k0, _, _ = _condense(lib, sd0, S0, K, empty, ctrl, lin0, np.zeros((1, len(ctrl), S0.c_stride)))
    assert np.allclose(kq[0, i], k0[0, i], rtol=1e-11, atol=1e-12)


def test_expand_update_consistency():
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(False)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=2, seed=11)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    dims = ANYMAL
    kk, ric, d, info = oracle_lib.riccati_batch(dims, K, ctrl, kkt, dx0)
    assert info == 0
    batch, n_grid = 2, len(ctrl)
    xd = np.zeros((batch, n_grid, S.x_stride))
    steps = np.zeros((batch, 2))
    csd = sd.c()
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(lin), oracle_lib.ptr(ex),
                         oracle_lib.ptr(d), oracle_lib.ptr(cc), oracle_lib.ptr(xd), oracle_lib.ptr(steps), 1)
    assert np.all(steps > 0) and np.all(steps <= 1)
    nv, nu, nx, npass, nvfm = 18, 12, 36, 6, 30
    for b in range(batch):
        for i in range(n_grid - 1):
            c = ctrl[i]
            nf = c.nf
            l, e = lin[b, i], ex[b, i]
            M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, 12, nv)[:nf]
            D, IDC = mat(l, S.l_D, nvfm, nx)[:nv + nf], l[S.l_IDC:S.l_IDC + nv + nf]
            dx, du = d[b, i, K.d_dx:K.d_dx + nx], d[b, i, K.d_du:K.d_du + nu]
            daf = xd[b, i, S.x_daf:S.x_daf + nv + nf].copy()
            daf[nv:] *= -1.0  # the reference flips the sign of df after the solve
            rhs = -D @ dx - IDC
            if c.type != IMPACT:
                rhs[npass:nv] += du
            full = np.block([[M, J.T], [J, np.zeros((nf, nf))]])
            assert np.allclose(full @ daf, rhs, rtol=1e-9, atol=1e-10)
    # fraction-to-boundary: the stepped slack / dual stay positive on active rows
    sl, dsl = cc[..., S.c_slack:S.c_slack + S.nc], cc[..., S.c_dslack:S.c_dslack + S.nc]
    du_, ddu = cc[..., S.c_dual:S.c_dual + S.nc], cc[..., S.c_ddual:S.c_ddual + S.nc]
    inter = np.array([c.type not in (IMPACT, TERMINAL) for c in ctrl])
    for b in range(batch):
        assert np.all((sl[b] + steps[b, 0] * dsl[b])[inter] > 0)
        assert np.all((du_[b] + steps[b, 1] * ddu[b])[inter] > 0)
    sol2, cc2, d2, ex2 = sol.copy(), cc.copy(), d.copy(), ex.copy()
    lib.orc_update_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(ex2), oracle_lib.ptr(d2),
                         oracle_lib.ptr(xd), oracle_lib.ptr(cc2), oracle_lib.ptr(sol2), oracle_lib.ptr(steps), 1)
    for b in range(batch):
        a = steps[b, 0]
        assert np.allclose(sol2[b, :, S.s_v:S.s_v + nv], sol[b, :, S.s_v:S.s_v + nv] + a * d[b, :, K.d_dx + nv:K.d_dx + nx])
        quat = sol2[b, :, S.s_q + 3:S.s_q + 7]
        assert np.allclose(np.linalg.norm(quat, axis=1), 1.0)
        assert np.allclose(sol2[b, :, S.s_q + 7:S.s_q + 19], sol[b, :, S.s_q + 7:S.s_q + 19] + a * d[b, :, K.d_dx + 6:K.d_dx + nv])
        assert np.allclose(cc2[b][inter][:, S.c_slack:S.c_slack + S.nc], (sl[b] + a * dsl[b])[inter])


@pytest.mark.parametrize("sto", [False, True])
def test_dual_expansion_satisfies_uncondensed_stationarity(sto):
    """expandContactDynamicsDual (contact_dynamics.cpp:177-202) pinned without its formulas: the multiplier directions it
    returns make the UNCONDENSED stage Lagrangian stationary in the eliminated variables,
        d/da : la' + Qaa' da + M dbeta + J^T dmu + dt dgmm+ (+ Phia^T dxi)            = 0
        d/df : lf' + Qff' df + Qqf'^T dq - J dbeta                                      = 0
        d/du : lu' + Quu' du - dbeta[actuated]                                          = 0
      (with switching-time optimisation each gets + dts * (ha | hf | hu), dts = (dts_next - dts) / num_grids_in_phase)
        passive joints : lu_passive - dbeta[passive] + dnu_passive                      = 0   (*)
    where ' marks the cost terms plus the PDIPM terms of the inequality rows (recomputed here as J^T diag(z/s) J and
    J^T cond).  Intermediate and lift stages incl. the switching-constraint stage of an ANYmal schedule.
    (*) on a switching-constraint stage the reference's dnu_passive (contact_dynamics.cpp:182-189) has no dxi term although
    dbeta does (:193-195), so there the residual is Z[passive, a] Phia^T dxi instead of 0 -- restated, and asserted as such;
    the same holds for the dts * haf term (:197-200)."""
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(sto)  # sto: the stage terms get + dts * (ha, hf), dts = (dts_next - dts) / N_phase
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=2, seed=13)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    # with STO the Riccati step is made the exact Newton step (test-only switches, see test_oracle_kkt.py): the reference's
    # own approximations in the phase transition would otherwise show up as stationarity residuals next to a transition
    flags = [ctypes.c_int.in_dll(lib, n) for n in ("orc_debug_exact_chi", "orc_debug_exact_transition", "orc_debug_exact_impact_costate")]
    try:
        for f in flags:
            f.value = int(sto)
        kk, ric, d, info = oracle_lib.riccati_batch(ANYMAL, K, ctrl, kkt, dx0, max_dts0=1e9 if sto else 0.1)
    finally:
        for f in flags:
            f.value = 0
    assert info == 0
    batch, n_grid = 2, len(ctrl)
    xd, steps = np.zeros((batch, n_grid, S.x_stride)), np.zeros((batch, 2))
    csd = sd.c()
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(lin), oracle_lib.ptr(ex),
                         oracle_lib.ptr(d), oracle_lib.ptr(cc), oracle_lib.ptr(xd), oracle_lib.ptr(steps), 1)
    sol2, cc2, d2, ex2 = sol.copy(), cc.copy(), d.copy(), ex.copy()
    lib.orc_update_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(ex2), oracle_lib.ptr(d2),
                         oracle_lib.ptr(xd), oracle_lib.ptr(cc2), oracle_lib.ptr(sol2), oracle_lib.ptr(steps), 1)
    nv, nu, nx, npass = 18, 12, 36, 6
    checked = 0
    for b in range(batch):
        for i in range(n_grid - 1):
            c = ctrl[i]
            if c.type == IMPACT:
                # expandImpactDynamicsDual (impact_dynamics.cpp:90-96): the impulse change ddv plays the role of da, no dt
                nf, l, x = c.nf, lin[b, i], xd[b, i]
                M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, 12, nv)[:nf]
                ddv, df = x[S.x_daf:S.x_daf + nv], x[S.x_daf + nv:S.x_daf + nv + nf]
                dbeta, dmu = x[S.x_dbetamu:S.x_dbetamu + nv], x[S.x_dbetamu + nv:S.x_dbetamu + nv + nf]
                dgmm_n = d[b, i + 1, K.d_dlmdgmm + nv:K.d_dlmdgmm + nx]
                ra = l[S.l_la:S.l_la + nv] + l[S.l_Qaa:S.l_Qaa + nv] * ddv + M @ dbeta + J.T @ dmu + dgmm_n
                assert np.abs(ra).max() < 1e-9 * max(np.abs(M @ dbeta).max(), 1.0), f"impact d/ddv stage {i}"
                rf = (l[S.l_lf:S.l_lf + nf] + mat(l, S.l_Qff, 12, 12)[:nf, :nf] @ df
                      + mat(l, S.l_Qqf, nv, 12)[:, :nf].T @ d[b, i, K.d_dx:K.d_dx + nv] - J @ dbeta)
                assert np.abs(rf).max() < 1e-9 * max(np.abs(J @ dbeta).max(), 1.0), f"impact d/df stage {i}"
                checked += 1
                continue
            nf, dt, l = c.nf, c.dt, lin[b, i]
            M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, 12, nv)[:nf]
            Qaa = l[S.l_Qaa:S.l_Qaa + nv].copy()
            Qff, Qqf = mat(l, S.l_Qff, 12, 12)[:nf, :nf].copy(), mat(l, S.l_Qqf, nv, 12)[:, :nf].copy()
            Quu = mat(l, S.l_Quu, nu, nu).copy()
            la, lf, lu = l[S.l_la:S.l_la + nv].copy(), l[S.l_lf:S.l_lf + nf].copy(), l[S.l_lu:S.l_lu + nu].copy()
            lup = l[S.l_lup:S.l_lup + npass]
            # PDIPM terms, recomputed from slack / dual / residual
            sl, du_, rs = (con[b, i, o:o + S.nc] for o in (S.c_slack, S.c_dual, S.c_res))
            w = du_ / sl
            cond = (du_ * rs - (sl * du_ - table.barrier)) / sl
            for r in range(table.n_box):
                br = table.box[r]
                if br.var == 3:  # torque limits
                    Quu[br.idx, br.idx] += w[r]
                    lu[br.idx] += br.sign * cond[r]
                elif br.var == 2:
                    Qaa[br.idx] += w[r]
                    la[br.idx] += br.sign * cond[r]
            fst = 0
            for ci in range(table.n_contacts):
                if not (c.contact_mask >> ci) & 1:
                    continue
                rows = slice(table.n_box + 5 * ci, table.n_box + 5 * ci + 5)
                dgdq = mat(l, S.l_dgdq + ci * 5 * nv, 5, nv)
                dgdf = mat(l, S.l_dgdf + ci * 15, 5, 3)
                Qff[fst:fst + 3, fst:fst + 3] += dgdf.T @ np.diag(w[rows]) @ dgdf
                Qqf[:, fst:fst + 3] += dgdq.T @ np.diag(w[rows]) @ dgdf
                lf[fst:fst + 3] += dgdf.T @ cond[rows]
                fst += 3
            di, dn, x = d[b, i], d[b, i + 1], xd[b, i]
            dq, du = di[K.d_dx:K.d_dx + nv], di[K.d_du:K.d_du + nu]
            da, df = x[S.x_daf:S.x_daf + nv], x[S.x_daf + nv:S.x_daf + nv + nf]
            dbeta, dmu = x[S.x_dbetamu:S.x_dbetamu + nv], x[S.x_dbetamu + nv:S.x_dbetamu + nv + nf]
            dnup = x[S.x_dnup:S.x_dnup + npass]
            dgmm_n = dn[K.d_dlmdgmm + nv:K.d_dlmdgmm + nx]
            dts = (di[K.d_dts + 1] - di[K.d_dts]) / c.ngrids_in_phase if c.sto else 0.0   # intermediate_stage.cpp:167-170
            ra = la + Qaa * da + M @ dbeta + J.T @ dmu + dt * dgmm_n + dts * l[S.l_ha:S.l_ha + nv]
            lf = lf + dts * l[S.l_hf:S.l_hf + nf]
            if c.ns > 0:
                ra = ra + mat(l, S.l_Phia, c.ns, nv).T @ di[K.d_dxi:K.d_dxi + c.ns]
            scale = max(np.abs(la).max(), np.abs(M @ dbeta).max(), 1.0)
            assert np.abs(ra).max() < 1e-9 * scale, f"d/da stage {i}"
            assert np.abs(lf + Qff @ df + Qqf.T @ dq - J @ dbeta).max() < 1e-9 * max(np.abs(J @ dbeta).max(), 1.0), f"d/df stage {i}"
            ru = lu + Quu @ du - dbeta[npass:] + dts * l[S.l_hu:S.l_hu + nu]
            assert np.abs(ru).max() < 1e-9 * max(np.abs(dbeta).max(), 1.0), f"d/du stage {i}"
            rp = lup - dbeta[:npass] + dnup
            Zfull = np.linalg.inv(np.block([[M, J.T], [J, np.zeros((nf, nf))]]))
            if c.ns > 0:
                rp = rp - Zfull[:npass, :nv] @ (mat(l, S.l_Phia, c.ns, nv).T @ di[K.d_dxi:K.d_dxi + c.ns])
            if dts != 0.0:  # (*) likewise the reference's dnu_passive has no dts * haf term
                rp = rp - Zfull[:npass, :] @ (dts * np.concatenate([l[S.l_ha:S.l_ha + nv], -l[S.l_hf:S.l_hf + nf]]))
            assert np.abs(rp).max() < 1e-9 * max(np.abs(dbeta).max(), 1.0), f"passive stage {i}"
            checked += 1
    assert checked == 2 * (n_grid - 1) and any(c.ns > 0 for c in ctrl) and any(c.type == IMPACT for c in ctrl)


def test_floating_base_state_equation_correction():
    """correctLinearizeStateEquation / correctCostateDirection (state_equation.cpp:68-95, se3_jacobian_inverse.hxx:17-32)
    pinned without the block-inverse formula: with E = dSubtract/dq+ (the 6x6 block the reference inverts) the corrected
    blocks satisfy  E Fqq[0:6,0:6] = -dSub/dqf,  E Fqv[0:6,0:6] = -dt I,  E Fq[0:6] = -Fq_uncorrected[0:6], and the costate
    head satisfies  Fqq_prev^T dlmd_corrected[0:6] = -dlmd[0:6]."""
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(False)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=1, seed=17)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    kkt0 = kkt.copy()
    kk, ric, d, info = oracle_lib.riccati_batch(ANYMAL, K, ctrl, kkt, dx0)
    n_grid = len(ctrl)
    xd, steps = np.zeros((1, n_grid, S.x_stride)), np.zeros((1, 2))
    csd = sd.c()
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, 1, oracle_lib.ptr(lin), oracle_lib.ptr(ex),
                         oracle_lib.ptr(d), oracle_lib.ptr(cc), oracle_lib.ptr(xd), oracle_lib.ptr(steps), 1)
    d_before = d.copy()
    sol2 = sol.copy()
    lib.orc_update_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, 1, oracle_lib.ptr(ex), oracle_lib.ptr(d),
                         oracle_lib.ptr(xd), oracle_lib.ptr(cc), oracle_lib.ptr(sol2), oracle_lib.ptr(steps), 1)
    nv, nx = 18, 36
    for i in range(n_grid):
        l, k = lin[0, i], kkt0[0, i]
        Fqq_prev = mat(l, S.l_se3 + 36, 6, 6)
        old, new = d_before[0, i, K.d_dlmdgmm:K.d_dlmdgmm + 6], d[0, i, K.d_dlmdgmm:K.d_dlmdgmm + 6]
        assert np.allclose(Fqq_prev.T @ new, -old, rtol=1e-10, atol=1e-12), f"costate head, grid {i}"
        assert np.array_equal(d[0, i, K.d_dlmdgmm + 6:K.d_dlmdgmm + nx], d_before[0, i, K.d_dlmdgmm + 6:K.d_dlmdgmm + nx])
        if ctrl[i].type == TERMINAL:
            continue
        E, dSub = mat(l, S.l_se3 + 72, 6, 6), mat(l, S.l_se3, 6, 6)
        Fxx = mat(k, K.k_Fxx, nx, nx)
        assert np.allclose(E @ Fxx[:6, :6], -dSub, rtol=1e-10, atol=1e-12), f"Fqq, grid {i}"
        if ctrl[i].type == IMPACT:
            assert not Fxx[:6, nv:nv + 6].any()
        else:
            assert np.allclose(E @ Fxx[:6, nv:nv + 6], -ctrl[i].dt * np.eye(6), rtol=1e-10, atol=1e-12), f"Fqv, grid {i}"
        assert np.allclose(E @ k[K.k_Fx:K.k_Fx + 6], -l[S.l_Fx:S.l_Fx + 6], rtol=1e-10, atol=1e-12), f"Fq, grid {i}"
        # the rest of the q rows is the fixed-base structure: identity / dt identity  (state_equation.cpp:52-55)
        assert np.array_equal(Fxx[6:nv, :nv], np.eye(nv)[6:]) and np.array_equal(Fxx[:6, 6:nv], np.zeros((6, nv - 6)))


def test_sto_sensitivities_are_the_substituted_hamiltonian_derivatives():
    """STO sensitivities of the condensing (contact_dynamics.cpp:155-163) + the 1/num_grids_in_phase scaling
    (intermediate_stage.cpp:140-148): with (a, f) = -R dx + Z[:, u] du - r substituted into the time derivative of the stage
    Lagrangian  h + hx^T dx + hu^T du + ha^T da + hf^T df,  the coefficients are
        hx - R^T [ha; -hf],   hu + Z[u, :] [ha; -hf],   h - r^T [ha; -hf]      (no inequality rows: the Qqf term vanishes)
    with R = Z dIDCdqv, r = Z IDC and Z recomputed here as a dense inverse."""
    table = rbt_constraint_table()
    table.barrier, table.fraction_to_boundary = 1e-3, 0.995
    td, ev, ctrl = small_event_schedule(True)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=1, seed=19)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    nv, nu, nx, npass, nvfm = 18, 12, 36, 6, 30
    checked = 0
    for i, c in enumerate(ctrl):
        if c.type in (IMPACT, TERMINAL) or not c.sto:
            continue
        l, k, nf, g1 = lin[0, i], kkt[0, i], c.nf, 1.0 / c.ngrids_in_phase
        M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, 12, nv)[:nf]
        Z = np.linalg.inv(np.block([[M, J.T], [J, np.zeros((nf, nf))]]))
        D, IDC = mat(l, S.l_D, nvfm, nx)[:nv + nf], l[S.l_IDC:S.l_IDC + nv + nf]
        R, r = Z @ D, Z @ IDC
        haf = np.concatenate([l[S.l_ha:S.l_ha + nv], -l[S.l_hf:S.l_hf + nf]])
        hx = (l[S.l_hx:S.l_hx + nx] - R.T @ haf) * g1
        hu = (l[S.l_hu:S.l_hu + nu] + Z[npass:npass + nu, :] @ haf) * g1
        h = (l[S.l_sc + 0] - r @ haf) * g1
        assert np.allclose(k[K.k_hx:K.k_hx + nx], hx, rtol=1e-10, atol=1e-12), f"hx, grid {i}"
        assert np.allclose(k[K.k_hu:K.k_hu + nu], hu, rtol=1e-10, atol=1e-12), f"hu, grid {i}"
        assert np.isclose(k[K.k_sc + 2], h, rtol=1e-10, atol=1e-12), f"h, grid {i}"
        assert np.isclose(k[K.k_sc + 0], l[S.l_sc + 1] * g1 * g1) and np.isclose(k[K.k_sc + 1], -k[K.k_sc + 0])
        fx = l[S.l_fx:S.l_fx + nx].copy()  # fq head corrected on SE(3): E fq' = -fq  (state_equation.cpp:85)
        E = mat(l, S.l_se3 + 72, 6, 6)
        got = k[K.k_fx:K.k_fx + nx] / g1
        assert np.allclose(E @ got[:6], -fx[:6], rtol=1e-10, atol=1e-12) and np.allclose(got[6:], fx[6:], rtol=1e-12, atol=0)
        checked += 1
    assert checked >= 8


def test_switching_constraint_condensing_is_the_substituted_constraint():
    """contact_dynamics.cpp:138-153: with da = -R_a dx + Z_au du - r_a substituted into  Phix dx + Phia da + p (+ Phit dt'),
    the condensed blocks are  Phix - Phia R_a,  Phia Z[a, u],  p - Phia r_a,  (Phit - Phia r_a) / num_grids_in_phase."""
    table = rbt_constraint_table()
    table.barrier, table.fraction_to_boundary = 1e-3, 0.995
    td, ev, ctrl = small_event_schedule(True)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=1, seed=23)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    nv, nu, nx, npass, nvfm = 18, 12, 36, 6, 30
    seen = 0
    for i, c in enumerate(ctrl):
        if c.ns == 0 or c.type in (IMPACT, TERMINAL):
            continue
        l, k, nf, ns = lin[0, i], kkt[0, i], c.nf, c.ns
        M, J = mat(l, S.l_M, nv, nv), mat(l, S.l_J, 12, nv)[:nf]
        Z = np.linalg.inv(np.block([[M, J.T], [J, np.zeros((nf, nf))]]))
        R = Z @ mat(l, S.l_D, nvfm, nx)[:nv + nf]
        r = Z @ l[S.l_IDC:S.l_IDC + nv + nf]
        Phia = mat(l, S.l_Phia, ns, nv)
        assert np.allclose(mat(k, K.k_Phix, ns, nx), mat(l, S.l_Phix, ns, nx) - Phia @ R[:nv], rtol=1e-10, atol=1e-12)
        assert np.allclose(mat(k, K.k_Phiu, ns, nu), Phia @ Z[:nv, npass:npass + nu], rtol=1e-10, atol=1e-12)
        assert np.allclose(k[K.k_p:K.k_p + ns], l[S.l_p:S.l_p + ns] - Phia @ r[:nv], rtol=1e-10, atol=1e-12)
        assert np.allclose(k[K.k_Phit:K.k_Phit + ns], (l[S.l_Phit:S.l_Phit + ns] - Phia @ r[:nv]) / c.ngrids_in_phase,
                           rtol=1e-10, atol=1e-12)
        seen += 1
    assert seen == 1


def test_constraint_expansion_is_the_linearised_pdipm_system():
    """expandSlackAndDual (joint_*_limit.cpp:78-83, friction_cone.cpp:238-268, pdipm.hxx:159-164): for every active row
    g'(x) d + dslack + residual = 0  and  slack*ddual + dual*dslack + (slack*dual - barrier) = 0, with g' rebuilt here from the
    constraint table and the friction-cone Jacobians; inactive cones get (1, 1) and never bind the step size."""
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(False)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=1, seed=29)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    kk, ric, d, info = oracle_lib.riccati_batch(ANYMAL, K, ctrl, kkt, dx0)
    n_grid = len(ctrl)
    xd, steps = np.zeros((1, n_grid, S.x_stride)), np.zeros((1, 2))
    csd = sd.c()
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, 1, oracle_lib.ptr(lin), oracle_lib.ptr(ex),
                         oracle_lib.ptr(d), oracle_lib.ptr(cc), oracle_lib.ptr(xd), oracle_lib.ptr(steps), 1)
    nv, nu = 18, 12
    lo_p, lo_d = 1.0, 1.0
    for i, c in enumerate(ctrl):
        if c.type in (IMPACT, TERMINAL):
            continue
        l, di, x, cr = lin[0, i], d[0, i], xd[0, i], cc[0, i]
        sl, du_, rs, cm, ds, dd = (cr[o:o + S.nc] for o in (S.c_slack, S.c_dual, S.c_res, S.c_cmpl, S.c_dslack, S.c_ddual))
        var = {0: di[K.d_dx:K.d_dx + nv], 1: di[K.d_dx + nv:K.d_dx + 2 * nv], 2: x[S.x_daf:S.x_daf + nv], 3: di[K.d_du:K.d_du + nu]}
        gd = np.zeros(S.nc)
        active = np.ones(S.nc, bool)
        for r in range(table.n_box):
            gd[r] = table.box[r].sign * var[table.box[r].var][table.box[r].idx]
            if {0: 2, 1: 1}.get(table.box[r].var, 0) + c.ineq_gate > 2:
                active[r] = False  # position- / velocity-level limits do not act on the first two grid points
        fst = 0
        for ci in range(table.n_contacts):
            rows = slice(table.n_box + 5 * ci, table.n_box + 5 * ci + 5)
            if not (c.contact_mask >> ci) & 1:
                active[rows] = False
                assert np.all(ds[rows] == 1.0) and np.all(dd[rows] == 1.0)
                continue
            gd[rows] = mat(l, S.l_dgdq + ci * 5 * nv, 5, nv) @ var[0] + mat(l, S.l_dgdf + ci * 15, 5, 3) @ x[S.x_daf + nv + fst:S.x_daf + nv + fst + 3]
            fst += 3
        assert np.allclose((gd + ds + rs)[active], 0, atol=1e-11)
        assert np.allclose((sl * dd + du_ * ds + (sl * du_ - table.barrier))[active], 0, atol=1e-11)
        assert np.allclose(cm[active], (sl * du_ - table.barrier)[active], rtol=1e-13)
        on = np.ones(S.nc, bool)
        on[:table.n_box] = active[:table.n_box]  # gated rows take no part in the step sizes (their dslack / ddual are untouched)
        with np.errstate(divide="ignore", invalid="ignore"):
            fp, fd = -table.fraction_to_boundary * (sl / ds), -table.fraction_to_boundary * (du_ / dd)
        lo_p = min([lo_p] + [f for f in fp[on] if 0 < f < 1])
        lo_d = min([lo_d] + [f for f in fd[on] if 0 < f < 1])
    assert steps[0, 0] == lo_p and steps[0, 1] == lo_d


def test_free_flyer_integration_is_the_se3_exponential():
    """SplitSolution::integrate on the floating base (robot.integrateConfiguration, Pinocchio's free-flyer joint:
    q (+) v = q * exp6(v)), restated in oracle/condense_oracle.c without Pinocchio -- checked here against an independent
    construction with scipy's Rotation: R+ = R exp(w^),  p+ = p + R V(w) v,  V = I + (1-cos t)/t^2 w^ + (t - sin t)/t^3 w^^2."""
    from scipy.spatial.transform import Rotation
    table = anymal_constraint_table()
    td, ev, ctrl = small_event_schedule(False)
    lib, sd, S, K, lin, con, sol, dx0 = _setup(table, ctrl, batch=2, seed=37)
    kkt, ex, cc = _condense(lib, sd, S, K, table, ctrl, lin, con)
    kk, ric, d, info = oracle_lib.riccati_batch(ANYMAL, K, ctrl, kkt, dx0)
    batch, n_grid = 2, len(ctrl)
    xd, steps = np.zeros((batch, n_grid, S.x_stride)), np.zeros((batch, 2))
    csd = sd.c()
    lib.orc_expand_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(lin), oracle_lib.ptr(ex),
                         oracle_lib.ptr(d), oracle_lib.ptr(cc), oracle_lib.ptr(xd), oracle_lib.ptr(steps), 1)
    d[:, :, K.d_dx:K.d_dx + 6] *= 25.0  # make the rotation large enough to exercise the trigonometric branch
    sol2 = sol.copy()
    lib.orc_update_batch(ctypes.byref(csd), ctypes.byref(table), ctrl, n_grid, batch, oracle_lib.ptr(ex), oracle_lib.ptr(d.copy()),
                         oracle_lib.ptr(xd), oracle_lib.ptr(cc), oracle_lib.ptr(sol2), oracle_lib.ptr(steps), 1)

    def hat(w):
        return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])

    biggest = 0.0
    for b in range(batch):
        a = steps[b, 0]
        for i in range(n_grid):
            q0, q1 = sol[b, i, S.s_q:S.s_q + 7], sol2[b, i, S.s_q:S.s_q + 7]
            v, w = a * d[b, i, K.d_dx:K.d_dx + 3], a * d[b, i, K.d_dx + 3:K.d_dx + 6]
            t = np.linalg.norm(w)
            biggest = max(biggest, t)
            W = hat(w)
            V = np.eye(3) + (1 - np.cos(t)) / t**2 * W + (t - np.sin(t)) / t**3 * W @ W
            R0 = Rotation.from_quat(q0[3:7])  # (x, y, z, w), Pinocchio's order
            p1 = q0[:3] + R0.apply(V @ v)
            R1 = R0 * Rotation.from_rotvec(w)
            assert np.allclose(q1[:3], p1, rtol=1e-11, atol=1e-12)
            assert np.allclose((Rotation.from_quat(q1[3:7]) * R1.inv()).magnitude(), 0.0, atol=1e-11)
    assert biggest > 0.05
