"""GPU parity: the CUDA path (through the C ABI / host mirror) against the CPU oracle on identical seeded inputs.
Tolerance: BASELINE.json north_star -- 1e-6 relative on P, K and the Newton direction (we assert 1e-8)."""
import numpy as np
import pytest

import oracle_lib
from helpers import jump_sto_schedule, rel_err, small_event_schedule, trot_schedule
from robotoc_b200 import ANYMAL, Layout, RiccatiRecursion, ULayout, UnconstrRiccatiRecursion
from robotoc_b200.grid import IMPACT
from synth import make_kkt, make_unconstr_kkt

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _blocks_ric(L, d, c):
    nx, nu, ns = d.nx, d.nu, c.ns
    b = {"P": (L.r_P, nx * nx), "s": (L.r_s, nx)}
    if c.type != IMPACT and c.type != 3:
        b.update({"K": (L.r_K, nx * nu), "k": (L.r_k, nu)})
        if ns > 0:
            b.update({"M": (L.r_M, ns * nx), "m": (L.r_m, ns)})
    if c.sto:
        b.update({"Psi": (L.r_Psi, nx), "Phi": (L.r_Phi, nx), "sc": (L.r_sc, 5)})
        if c.type != IMPACT:
            b.update({"T": (L.r_T, nu), "W": (L.r_W, nu), "psix": (L.r_psix, nx), "psiu": (L.r_psiu, nu),
                      "phix": (L.r_phix, nx), "phiu": (L.r_phiu, nu)})
            if ns > 0:
                b.update({"mt": (L.r_mt, ns), "mtn": (L.r_mtn, ns)})
    b.update({"dtsdx": (L.r_dtsdx, nx), "stosc": (L.r_stosc, 2)})
    return b


def _compare(dims, L, ctrl, got_ric, ref_ric, got_d, ref_d, got_f=None, ref_kkt=None, tol=TOL):
    worst = 0.0
    for i, c in enumerate(ctrl):
        for name, (off, n) in _blocks_ric(L, dims, c).items():
            a, b = got_ric[:, i, off:off + n], ref_ric[:, i, off:off + n]
            if np.max(np.abs(b)) == 0.0:
                assert np.max(np.abs(a)) == 0.0, f"stage {i} block {name}: expected zeros"
                continue
            if name == "W" and c.ns == dims.nu:
                # ns == nu: the projected inverse Ginv - SDG^T DG is analytically zero, so W = -Ginv phi_u is pure
                # cancellation noise (~1e-14); compare on the scale of its sibling T instead of its own.
                scale = np.max(np.abs(ref_ric[:, i, L.r_T:L.r_T + dims.nu]))
                assert np.max(np.abs(a - b)) < 1e-9 * scale, f"riccati stage {i} block W (ns==nu)"
                continue
            e = rel_err(a, b)
            worst = max(worst, e)
            assert e < tol, f"riccati stage {i} block {name}: rel err {e:.3e}"
        dblocks = {"dx": (L.d_dx, dims.nx), "dlmdgmm": (L.d_dlmdgmm, dims.nx), "dts": (L.d_dts, 2)}
        if c.type not in (IMPACT, 3):
            dblocks["du"] = (L.d_du, dims.nu)
            if c.ns > 0:
                dblocks["dxi"] = (L.d_dxi, c.ns)
        for name, (off, n) in dblocks.items():
            a, b = got_d[:, i, off:off + n], ref_d[:, i, off:off + n]
            if np.max(np.abs(b)) == 0.0:
                assert np.max(np.abs(a)) < 1e-300, f"stage {i} dir {name}: expected zeros"
                continue
            e = rel_err(a, b)
            worst = max(worst, e)
            assert e < tol, f"direction stage {i} block {name}: rel err {e:.3e}"
        if got_f is not None and c.type != 3:
            fb = {"F": (L.f_F, L.k_Qxx, dims.nx ** 2)}
            if c.type != IMPACT:
                fb.update({"H": (L.f_H, L.k_Qxu, dims.nx * dims.nu), "G": (L.f_G, L.k_Quu, dims.nu ** 2),
                           "lu": (L.f_lu, L.k_lu, dims.nu)})
            for name, (fo, ko, n) in fb.items():
                e = rel_err(got_f[:, i, fo:fo + n], ref_kkt[:, i, ko:ko + n])
                worst = max(worst, e)
                assert e < tol, f"factorized KKT stage {i} block {name}: rel err {e:.3e}"
    return worst


def _run_case(ctrl, batch, seed, max_dts0=0.1):
    dims = ANYMAL
    L = Layout(dims)
    kkt, dx0 = make_kkt(dims, L, ctrl, batch=batch, seed=seed)
    rr = RiccatiRecursion(dims, len(ctrl), batch, max_dts0=max_dts0)
    rr.setTimeDiscretization(ctrl)
    rr.backwardRiccatiRecursion(kkt, write_fact=True)
    rr.forwardRiccatiRecursion(dx0)
    ric, d, f = rr.getRiccatiFactorization(), rr.getDirection(), rr.getFactorizedKKT()
    assert int(rr.info().max()) == 0
    kk, ric_o, d_o, info = oracle_lib.riccati_batch(dims, L, ctrl, kkt, dx0, max_dts0=max_dts0)
    assert info == 0
    worst = _compare(dims, L, ctrl, ric, ric_o, d, d_o, f, kk)
    # one-call host API gives the same answer
    ric2, d2 = rr.solve_host(kkt, dx0)
    assert np.array_equal(ric2, ric) and np.array_equal(d2, d)
    rr.close()
    return worst


def test_small_event_schedule():
    td, ev, ctrl = small_event_schedule(sto=False)
    w = _run_case(ctrl, batch=5, seed=11)
    print("worst rel err", w)


def test_trot_n40_config3_small_batch():
    """BASELINE config 3 schedule (47 grid points: 2 Lift, 2 Impact, 2 switching stages), small batch."""
    td, ev, ctrl = trot_schedule(40)
    assert len(ctrl) == 47
    w = _run_case(ctrl, batch=6, seed=20260927)
    print("worst rel err", w)


def test_small_event_schedule_sto():
    td, ev, ctrl = small_event_schedule(sto=True)
    assert any(c.sto for c in ctrl)
    w = _run_case(ctrl, batch=4, seed=5)
    print("worst rel err", w)


def test_jump_sto_n80_config4_small_batch():
    """BASELINE config 4 schedule (84 grid points, all phases STO, ns=12 switching stage)."""
    td, ev, ctrl = jump_sto_schedule(80)
    assert len(ctrl) == 84
    w = _run_case(ctrl, batch=3, seed=20260928)
    print("worst rel err", w)


@pytest.mark.parametrize("N,batch,dt", [(20, 1, 0.05), (50, 16, 0.02)])
def test_unconstr_iiwa14(N, batch, dt):
    """BASELINE configs 1 and 2 (iiwa14, nv=7)."""
    nv = 7
    UL = ULayout(nv)
    kkt, dx0 = make_unconstr_kkt(nv, UL, N, batch, seed=20260925)
    ur = UnconstrRiccatiRecursion(nv, N, dt, batch)
    ur.backwardRiccatiRecursion(kkt, write_fact=True)
    ur.forwardRiccatiRecursion(dx0)
    ric, d, f = ur.getRiccatiFactorization(), ur.getDirection(), ur.getFactorizedKKT()
    assert int(ur.info().max()) == 0
    kk, ric_o, d_o, info = oracle_lib.unconstr_batch(nv, UL, N, dt, kkt, dx0)
    assert info == 0
    nx = 2 * nv
    for i in range(N + 1):
        blocks = {"P": (UL.r_P, nx * nx), "s": (UL.r_s, nx)}
        if i < N:
            blocks.update({"K": (UL.r_K, nx * nv), "k": (UL.r_k, nv)})
        for name, (off, n) in blocks.items():
            e = rel_err(ric[:, i, off:off + n], ric_o[:, i, off:off + n])
            assert e < TOL, f"stage {i} {name}: {e:.3e}"
        dbl = {"dx": (UL.d_dx, nx), "dlmdgmm": (UL.d_dlmdgmm, nx)}
        if i < N:
            dbl["da"] = (UL.d_da, nv)
        for name, (off, n) in dbl.items():
            e = rel_err(d[:, i, off:off + n], d_o[:, i, off:off + n])
            assert e < TOL, f"dir stage {i} {name}: {e:.3e}"
        if i < N:
            for name, fo, ko, n in [("F", UL.f_F, UL.k_Qxx, nx * nx), ("H", UL.f_H, UL.k_Qxu, nx * nv),
                                    ("G", UL.f_G, UL.k_Qaa, nv * nv), ("la", UL.f_la, UL.k_la, nv)]:
                e = rel_err(f[:, i, fo:fo + n], kk[:, i, ko:ko + n])
                assert e < TOL, f"fact stage {i} {name}: {e:.3e}"
    ric2, d2 = ur.solve_host(kkt, dx0)
    assert np.array_equal(ric2, ric) and np.array_equal(d2, d)
    ur.close()
