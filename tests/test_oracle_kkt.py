"""Pins the CPU oracle (oracle/riccati_oracle.c) independently of the Riccati algebra:
the Newton direction it returns must solve the full block-tridiagonal KKT system (numpy dense solve).
This covers Intermediate, Lift, Impact stages and the switching-constraint (Schur) path."""
import numpy as np
import pytest

import oracle_lib
from helpers import dense_kkt_solve, small_event_schedule, rel_err
from robotoc_b200 import ANYMAL, Layout
from robotoc_b200.schedule import IMPACT, LIFT
from robotoc_b200.synth import make_kkt


@pytest.mark.parametrize("seed", [20260924, 7])
def test_oracle_direction_solves_full_kkt(seed):
    dims = ANYMAL
    L = Layout(dims)
    td, ev, ctrl = small_event_schedule(sto=False)
    types = [c.type for c in ctrl]
    assert IMPACT in types and LIFT in types and any(c.ns > 0 for c in ctrl)
    kkt, dx0 = make_kkt(dims, L, ctrl, batch=2, seed=seed)
    kk, ric, d, info = oracle_lib.riccati_batch(dims, L, ctrl, kkt, dx0)
    assert info == 0
    for b in range(2):
        ref = dense_kkt_solve(dims, L, ctrl, kkt[b], dx0[b])
        for i in range(len(ctrl)):
            di = d[b, i]
            assert rel_err(di[L.d_dx:L.d_dx + dims.nx], ref[("dx", i)]) < 1e-9
            assert rel_err(di[L.d_dlmdgmm:L.d_dlmdgmm + dims.nx], ref[("lmd", i)]) < 1e-8
            if ("du", i) in ref:
                assert rel_err(di[L.d_du:L.d_du + dims.nu], ref[("du", i)]) < 1e-9
            if ("xi", i) in ref:
                ns = ctrl[i].ns
                assert rel_err(di[L.d_dxi:L.d_dxi + ns], ref[("xi", i)]) < 1e-8


def test_oracle_riccati_symmetry_and_mutation():
    """reference identities: P symmetric (backward_..factorizer.cpp:85); K = -G^-1 H^T on plain stages
    (riccati_factorizer.cpp:55); the mutated Qxx equals F - K^T G K (test/riccati/riccati_factorizer_test.cpp:36-71)."""
    dims = ANYMAL
    L = Layout(dims)
    td, ev, ctrl = small_event_schedule(sto=False)
    kkt, dx0 = make_kkt(dims, L, ctrl, batch=1, seed=3)
    kk, ric, d, info = oracle_lib.riccati_batch(dims, L, ctrl, kkt, dx0)
    nx, nu = dims.nx, dims.nu
    from robotoc_b200.synth import mat
    for i in range(len(ctrl) - 1):
        P = mat(ric[0, i], L.r_P, nx, nx)
        assert np.allclose(P, P.T, rtol=0, atol=1e-12 * np.abs(P).max())
        if ctrl[i].type != IMPACT and ctrl[i].ns == 0:
            G = mat(kk[0, i], L.k_Quu, nu, nu)
            H = mat(kk[0, i], L.k_Qxu, nx, nu)
            Kt = mat(ric[0, i], L.r_K, nx, nu)
            assert rel_err(Kt.T, -np.linalg.solve(G, H.T)) < 1e-10
