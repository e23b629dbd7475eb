"""Per-stage condense / expand / update layer: layouts, constraint table, synthetic inputs.

Mirrors, on the host side,
  the tail of IntermediateStage/ImpactStage/TerminalStage::evalKKT ("Forms linear system",
      /root/reference/src/ocp/intermediate_stage.cpp:133-148, impact_stage.cpp:115-121, terminal_stage.cpp:102-106),
  DirectMultipleShooting::computeStepSizes / maxPrimalStepSize / maxDualStepSize / integrateSolution
      (/root/reference/src/ocp/direct_multiple_shooting.cpp:174-241).
The records are described in include/rbt_stage_layout.h.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .layout import Dims
from .schedule import IMPACT, TERMINAL

RBT_MAX_BOX_ROWS = 128
VAR_Q, VAR_V, VAR_A, VAR_U = 0, 1, 2, 3


class rbt_box_row(ctypes.Structure):
    _fields_ = [("var", ctypes.c_int), ("idx", ctypes.c_int), ("sign", ctypes.c_int)]


class rbt_constraint_table(ctypes.Structure):
    _fields_ = [("n_box", ctypes.c_int), ("n_contacts", ctypes.c_int), ("barrier", ctypes.c_double),
                ("fraction_to_boundary", ctypes.c_double), ("box", rbt_box_row * RBT_MAX_BOX_ROWS)]


class rbt_stage_dims(ctypes.Structure):
    _fields_ = [("nv", ctypes.c_int), ("nu", ctypes.c_int), ("n_passive", ctypes.c_int), ("nf_max", ctypes.c_int),
                ("ns_max", ctypes.c_int), ("n_contacts", ctypes.c_int), ("n_box", ctypes.c_int)]


_SFIELDS = ("nv nu nx np nfm nvf nsm ncon nbox nc ncp nq l_M l_J l_D l_IDC l_Qaa l_Qff l_Qqf l_Qxx l_Quu l_lx l_la l_lf l_lu "
            "l_Fx l_lup l_se3 l_Phix l_Phia l_p l_Phit l_ha l_hf l_hx l_hu l_fx l_sc l_dgdq l_dgdf l_stride e_Z e_R e_r "
            "e_Qafqv e_Qafu e_laf e_Qxup e_Quup e_lup e_Phia e_haf e_Fqqpi e_stride c_slack c_dual c_res c_cmpl c_cond "
            "c_dslack c_ddual c_stride s_q s_v s_a s_dv s_u s_f s_lmd s_gmm s_beta s_mu s_nup s_xi s_stride x_daf "
            "x_dbetamu x_dnup x_stride").split()


def anymal_constraint_table(barrier=1.0e-3, fraction_to_boundary=0.995):
    """examples/anymal/trot.cpp:131-148: joint position / velocity / torque lower+upper limits on the 12 actuated joints
    (72 box rows) and friction cones on the 4 feet (5 rows each) -> 92 inequalities per stage."""
    t = rbt_constraint_table()
    t.n_contacts = 4
    t.barrier, t.fraction_to_boundary = barrier, fraction_to_boundary
    r = 0
    for var, off in ((VAR_Q, 6), (VAR_V, 6), (VAR_U, 0)):
        for sign in (-1, +1):
            for j in range(12):
                t.box[r].var, t.box[r].idx, t.box[r].sign = var, off + j, sign
                r += 1
    t.n_box = r
    return t


@dataclass(frozen=True)
class StageDims:
    dims: Dims
    nf_max: int
    n_contacts: int
    n_box: int

    def c(self):
        d = self.dims
        return rbt_stage_dims(d.nv, d.nu, d.n_passive, self.nf_max, d.ns_max, self.n_contacts, self.n_box)


class StageLayout:
    def __init__(self, sdims: StageDims, getter=None):
        self.sdims = sdims
        cd = sdims.c()
        get = getter or _lib.lib().rbt_stage_layout_get
        for f in _SFIELDS:
            v = get(ctypes.byref(cd), f.encode())
            if v < 0:
                raise RuntimeError(f"stage layout field {f} unknown to the library")
            setattr(self, f, v)


def _u(rng, *shape):
    return rng.uniform(-1.0, 1.0, size=shape)


def _putm(rec, off, block, ld):
    """block[b, i, j] -> rec[b, off + i + j*ld] (column-major with leading dimension ld)."""
    b, m, n = block.shape
    view = rec[:, off:off + ld * n].reshape(b, n, ld)
    view[:, :, :m] = np.transpose(block, (0, 2, 1))


def make_stage_inputs(sd: StageDims, S: StageLayout, ctrl, batch: int, seed: int):
    """Synthetic linearization / PDIPM / solution records (there is no Pinocchio here).  Structure follows what the
    reference's linearize* halves produce: M SPD (joint-space inertia), J a contact Jacobian, diagonal Qaa, friction-cone
    Jacobians per active contact, SE(3) blocks [[A,B],[0,D]] (se3_jacobian_inverse.hxx), slack/dual > 0."""
    rng = np.random.default_rng(seed)
    d = sd.dims
    nv, nu, nx, np_, nfm, nvfm = d.nv, d.nu, d.nx, d.n_passive, sd.nf_max, d.nv + sd.nf_max
    n_grid = len(ctrl)
    lin = np.zeros((batch, n_grid, S.l_stride))
    con = np.zeros((batch, n_grid, S.c_stride))
    sol = np.zeros((batch, n_grid, S.s_stride))
    for i in range(n_grid):
        c = ctrl[i]
        rec = lin[:, i, :]
        Ss = _u(rng, batch, nx, nx)
        _putm(rec, S.l_Qxx, Ss @ np.transpose(Ss, (0, 2, 1)) / nx + np.eye(nx)[None], nx)
        rec[:, S.l_lx:S.l_lx + nx] = _u(rng, batch, nx)
        if np_ == 6:
            for k in range(3):
                blk = np.zeros((batch, 6, 6))
                blk[:, :3, :3] = np.eye(3) + 0.1 * _u(rng, batch, 3, 3)
                blk[:, 3:, 3:] = np.eye(3) + 0.1 * _u(rng, batch, 3, 3)
                blk[:, :3, 3:] = 0.1 * _u(rng, batch, 3, 3)
                if k == 0:
                    blk = -blk  # dSubtract/dqf ~ -I near q_next ~ q
                _putm(rec, S.l_se3 + 36 * k, blk, 6)
        # solution (all stages)
        q = _u(rng, batch, S.nq)
        if np_ == 6:
            quat = _u(rng, batch, 4)
            q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
        sol[:, i, S.s_q:S.s_q + S.nq] = q
        for off, n in ((S.s_v, nv), (S.s_a, nv), (S.s_dv, nv), (S.s_u, nu), (S.s_f, nfm), (S.s_lmd, nv), (S.s_gmm, nv),
                       (S.s_beta, nv), (S.s_mu, nfm), (S.s_nup, np_), (S.s_xi, S.nsm)):
            sol[:, i, off:off + n] = _u(rng, batch, n)
        if c.type == TERMINAL:
            continue
        impact = c.type == IMPACT
        nf = c.nf
        Sm = _u(rng, batch, nv, nv)
        _putm(rec, S.l_M, np.eye(nv)[None] + 0.1 * Sm @ np.transpose(Sm, (0, 2, 1)), nv)
        if nf > 0:
            _putm(rec, S.l_J, _u(rng, batch, nf, nv), nfm)
        D = 0.5 * _u(rng, batch, nv + nf, nx)
        if impact:
            D[:, :nv, nv:] = 0.0  # dIDdv does not exist on an impact stage (impact_dynamics.cpp:44-52)
        _putm(rec, S.l_D, D, nvfm)
        rec[:, S.l_IDC:S.l_IDC + nv + nf] = 0.5 * _u(rng, batch, nv + nf)
        rec[:, S.l_Qaa:S.l_Qaa + nv] = rng.uniform(0.01, 1.0, size=(batch, nv))
        if nf > 0:
            _putm(rec, S.l_Qff, np.repeat(1e-3 * np.eye(nf)[None], batch, 0), nfm)
        rec[:, S.l_la:S.l_la + nv] = _u(rng, batch, nv)
        rec[:, S.l_lf:S.l_lf + nf] = _u(rng, batch, nf)
        rec[:, S.l_Fx:S.l_Fx + nx] = 0.1 * _u(rng, batch, nx)
        if not impact:
            Su = _u(rng, batch, nu, nu)
            _putm(rec, S.l_Quu, Su @ np.transpose(Su, (0, 2, 1)) / nu + 0.1 * np.eye(nu)[None], nu)
            rec[:, S.l_lu:S.l_lu + nu] = _u(rng, batch, nu)
            rec[:, S.l_lup:S.l_lup + np_] = _u(rng, batch, np_)
            if c.ns > 0:
                _putm(rec, S.l_Phix, _u(rng, batch, c.ns, nx), c.ns)
                _putm(rec, S.l_Phia, _u(rng, batch, c.ns, nv), c.ns)
                rec[:, S.l_p:S.l_p + c.ns] = _u(rng, batch, c.ns)
                rec[:, S.l_Phit:S.l_Phit + c.ns] = _u(rng, batch, c.ns)
            rec[:, S.l_ha:S.l_ha + nv] = _u(rng, batch, nv)
            rec[:, S.l_hf:S.l_hf + nf] = _u(rng, batch, nf)
            rec[:, S.l_hx:S.l_hx + nx] = _u(rng, batch, nx)
            rec[:, S.l_hu:S.l_hu + nu] = _u(rng, batch, nu)
            rec[:, S.l_fx:S.l_fx + nx] = _u(rng, batch, nx)
            rec[:, S.l_sc + 0] = _u(rng, batch)
            rec[:, S.l_sc + 1] = rng.uniform(0.5, 1.5, size=batch)
            for ci in range(sd.n_contacts):
                if (c.contact_mask >> ci) & 1:
                    _putm(rec, S.l_dgdq + ci * 5 * nv, 0.3 * _u(rng, batch, 5, nv), 5)
                    _putm(rec, S.l_dgdf + ci * 15, _u(rng, batch, 5, 3), 5)
            nc = S.nc
            con[:, i, S.c_slack:S.c_slack + nc] = rng.uniform(0.01, 1.0, size=(batch, nc))
            con[:, i, S.c_dual:S.c_dual + nc] = rng.uniform(0.01, 1.0, size=(batch, nc))
            con[:, i, S.c_res:S.c_res + nc] = 0.1 * _u(rng, batch, nc)
    dx0 = 0.1 * _u(rng, batch, nx)
    return np.ascontiguousarray(lin), np.ascontiguousarray(con), np.ascontiguousarray(sol), np.ascontiguousarray(dx0)


def symmetrize_lin(S: StageLayout, lin):
    """Makes the symmetric blocks of linearization records exactly symmetric (upper triangle authoritative), which is what
    the reference's containers hold (cost Hessians, joint-space inertia) and what the host wire format assumes."""
    out = lin.copy()
    flat = out.reshape(-1, out.shape[-1])
    for off, n in ((S.l_M, S.nv), (S.l_Qff, S.nfm), (S.l_Qxx, S.nx), (S.l_Quu, S.nu)):
        blk = flat[:, off:off + n * n].reshape(-1, n, n)  # [rec, col, row] (column-major)
        a = np.transpose(blk, (0, 2, 1))                  # a[rec, row, col]
        up = np.triu(a)
        sym = up + np.transpose(np.triu(a, 1), (0, 2, 1))
        flat[:, off:off + n * n] = np.transpose(sym, (0, 2, 1)).reshape(-1, n * n)
    return out
