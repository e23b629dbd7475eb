"""Synthetic, structure-faithful KKT inputs (there is no Pinocchio here to linearise a real robot).

Recipes follow the reference's own test factories with FIXED seeds:
  /root/reference/test/test_helper/kkt_factory.cpp:7-64        (CreateSplitKKTMatrix / Residual)
  /root/reference/test/riccati/riccati_factorizer_test.cpp:131-140 (switching-constraint blocks)
  /root/reference/test/riccati/unconstr_riccati_recursion_test.cpp:33-46 (unconstrained horizon)
"Random" is U(-1,1) like Eigen::Random.  Documented deviation: the velocity rows of Fxx are
`Fvq = dt*R, Fvv = I + dt*R, Fvu = dt*R` and the floating-base 6x6 blocks are `I + dt*R` / `dt*(I+R)`, which is the
structure condensing produces (contact_dynamics.cpp:130-134, state_equation.cpp:80-82) and keeps a 40-90 step
recursion well conditioned; `contractive=False` gives the raw unit-scale factory blocks.
"""
import numpy as np

from .layout import Dims, Layout, ULayout
from .schedule import IMPACT, TERMINAL


def _u(rng, *shape):
    return rng.uniform(-1.0, 1.0, size=shape)


def _put(rec, off, block):
    """Store a batch of column-major blocks: block[b, i, j] -> rec[b, off + i + j*rows]."""
    b = block.shape[0]
    if block.ndim == 2:
        rec[:, off:off + block.shape[1]] = block
    else:
        rec[:, off:off + block.shape[1] * block.shape[2]] = np.transpose(block, (0, 2, 1)).reshape(b, -1)


def make_kkt(dims: Dims, L: Layout, ctrl, batch: int, seed: int, contractive: bool = True):
    """Returns kkt[batch, n_grid, k_stride] (float64, C-contiguous) and dx0[batch, nx]."""
    rng = np.random.default_rng(seed)
    n_grid = len(ctrl)
    nv, nu, nx = dims.nv, dims.nu, dims.nx
    kkt = np.zeros((batch, n_grid, L.k_stride))
    fb = dims.n_passive == 6
    eye = np.eye(nv)[None]
    for i in range(n_grid):
        c = ctrl[i]
        rec = kkt[:, i, :]
        if c.type == TERMINAL:
            S = _u(rng, batch, nx, nx)
            _put(rec, L.k_Qxx, S @ np.transpose(S, (0, 2, 1)))
            _put(rec, L.k_lx, _u(rng, batch, nx))
            continue
        dt = c.dt if c.dt > 0 else 0.0
        sc = dt if contractive else 1.0
        Fxx = np.zeros((batch, nx, nx))
        Fqq = np.repeat(eye, batch, 0).copy()
        if c.type == IMPACT:
            Fqv = np.zeros((batch, nv, nv))
            if fb:
                R = _u(rng, batch, 6, 6)
                Fqq[:, :6, :6] = (np.eye(6)[None] + 0.02 * R) if contractive else R
            Fvq = _u(rng, batch, nv, nv) * (0.05 if contractive else 1.0)
            Fvv = _u(rng, batch, nv, nv) * (0.05 if contractive else 1.0) + (eye if contractive else 0.0)
        else:
            Fqv = np.repeat(dt * eye, batch, 0).copy()
            if fb:
                R1, R2 = _u(rng, batch, 6, 6), _u(rng, batch, 6, 6)
                Fqq[:, :6, :6] = (np.eye(6)[None] + dt * R1) if contractive else R1
                Fqv[:, :6, :6] = (dt * (np.eye(6)[None] + 0.5 * R2)) if contractive else R2
            Fvq = sc * _u(rng, batch, nv, nv)
            Fvv = sc * _u(rng, batch, nv, nv) + (eye if contractive else 0.0)
        Fxx[:, :nv, :nv], Fxx[:, :nv, nv:], Fxx[:, nv:, :nv], Fxx[:, nv:, nv:] = Fqq, Fqv, Fvq, Fvv
        _put(rec, L.k_Fxx, Fxx)
        _put(rec, L.k_Fx, _u(rng, batch, nx))
        _put(rec, L.k_lx, _u(rng, batch, nx))
        if c.type == IMPACT:
            S = _u(rng, batch, nx, nx)
            _put(rec, L.k_Qxx, S @ np.transpose(S, (0, 2, 1)))
        else:
            _put(rec, L.k_Fvu, sc * _u(rng, batch, nv, nu))
            Hs = _u(rng, batch, nx + nu, nx + nu)
            H = Hs @ np.transpose(Hs, (0, 2, 1))
            _put(rec, L.k_Qxx, H[:, :nx, :nx])
            _put(rec, L.k_Qxu, H[:, :nx, nx:])
            _put(rec, L.k_Quu, H[:, nx:, nx:])
            _put(rec, L.k_lu, _u(rng, batch, nu))
            if c.ns > 0:
                _put(rec, L.k_Phix, _u(rng, batch, c.ns, nx))
                _put(rec, L.k_Phiu, _u(rng, batch, c.ns, nu))
                _put(rec, L.k_p, _u(rng, batch, c.ns))
            if c.sto:
                ng = max(c.ngrids_in_phase, 1)
                _put(rec, L.k_fx, _u(rng, batch, nx) / ng)
                _put(rec, L.k_hx, _u(rng, batch, nx) / ng)
                _put(rec, L.k_hu, _u(rng, batch, nu) / ng)
                if c.ns > 0:
                    _put(rec, L.k_Phit, _u(rng, batch, c.ns) / ng)
                qtt = rng.uniform(0.5, 1.5, size=batch) / (ng * ng) + 1.0
                rec[:, L.k_sc + 0] = qtt
                rec[:, L.k_sc + 1] = -qtt          # intermediate_stage.cpp:145
                rec[:, L.k_sc + 2] = _u(rng, batch) / ng
    dx0 = _u(rng, batch, nx)
    return np.ascontiguousarray(kkt), np.ascontiguousarray(dx0)


def make_unconstr_kkt(nv: int, UL: ULayout, N: int, batch: int, seed: int):
    """unconstr_riccati_recursion_test.cpp:33-46: [Qxx Qxu; . Qaa] = G G^T, G in U(-1,1)^{3nv x 3nv}; Fx,lx,la random."""
    rng = np.random.default_rng(seed)
    nx = 2 * nv
    kkt = np.zeros((batch, N + 1, UL.k_stride))
    for i in range(N + 1):
        rec = kkt[:, i, :]
        Gs = _u(rng, batch, 3 * nv, 3 * nv)
        H = Gs @ np.transpose(Gs, (0, 2, 1))
        _put(rec, UL.k_Qxx, H[:, :nx, :nx])
        _put(rec, UL.k_lx, _u(rng, batch, nx))
        if i < N:
            _put(rec, UL.k_Qxu, H[:, :nx, nx:])
            _put(rec, UL.k_Qaa, H[:, nx:, nx:])
            _put(rec, UL.k_Fx, _u(rng, batch, nx))
            _put(rec, UL.k_la, _u(rng, batch, nv))
    dx0 = _u(rng, batch, nx)
    return np.ascontiguousarray(kkt), np.ascontiguousarray(dx0)


def mat(rec, off, rows, cols):
    """View helper: column-major block at `off` of a single record -> (rows, cols) array."""
    return rec[off:off + rows * cols].reshape(cols, rows).T
