"""Error paths through the C ABI on a CUDA device: argument validation of the control table, call-order errors, and the
numerical-failure path (non-positive-definite Quu + B^T P B -> per-OCP flag, RBT_ERR_NUMERIC; the reference asserts
llt_.info() == Eigen::Success, riccati_factorizer.cpp:50)."""
import copy
import ctypes

import numpy as np
import pytest

import oracle_lib
from helpers import rel_err, small_event_schedule, trot_schedule
from robotoc_b200 import ANYMAL, DirectMultipleShooting, Layout, RiccatiRecursion, StageDims, anymal_constraint_table
from robotoc_b200._lib import rbt_stage_ctrl
from synth import make_kkt

pytestmark = pytest.mark.gpu


def _clone(ctrl):
    out = (rbt_stage_ctrl * len(ctrl))()
    for i, c in enumerate(ctrl):
        ctypes.memmove(ctypes.byref(out[i]), ctypes.byref(c), ctypes.sizeof(rbt_stage_ctrl))
    return out


@pytest.mark.parametrize("field,value", [("nf", 13), ("nf", -3), ("nf", 15), ("contact_mask", 0b0001), ("contact_mask", 1 << 9),
                                          ("ngrids_in_phase", -1), ("dt", -0.01), ("dt", float("nan")), ("dt", float("inf")),
                                          ("sto", 2), ("ns", 13)])
def test_set_schedule_rejects_inconsistent_contact_bookkeeping(field, value):
    td, ev, ctrl = small_event_schedule(False)
    rr = RiccatiRecursion(ANYMAL, len(ctrl), 2)
    rr.setTimeDiscretization(ctrl)  # the good table is accepted
    bad = _clone(ctrl)
    setattr(bad[2], field, value)
    with pytest.raises(ValueError):
        rr.setTimeDiscretization(bad)
    rr.close()


def test_impact_grid_cannot_carry_a_switching_constraint_and_stage_setup_is_not_repeatable():
    td, ev, ctrl = small_event_schedule(False)
    rr = RiccatiRecursion(ANYMAL, len(ctrl), 2)
    bad = _clone(ctrl)
    imp = [i for i, c in enumerate(ctrl) if c.type == 1][0]
    bad[imp].ns = 6
    with pytest.raises(ValueError):
        rr.setTimeDiscretization(bad)
    rr.setTimeDiscretization(ctrl)
    table = anymal_constraint_table()
    sd = StageDims(ANYMAL, nf_max=12, n_contacts=table.n_contacts, n_box=table.n_box)
    DirectMultipleShooting(rr, sd, table)
    with pytest.raises(RuntimeError):  # RBT_ERR_STATE: a second set-up would leak the stage buffers
        DirectMultipleShooting(rr, sd, table)
    rr.close()


def test_non_positive_definite_quu_flags_only_that_ocp():
    td, ev, ctrl = trot_schedule(40)
    dims, L = ANYMAL, Layout(ANYMAL)
    batch, sick, stage = 8, 5, 17
    kkt, dx0 = make_kkt(dims, L, ctrl, batch=batch, seed=77)
    assert ctrl[stage].type == 0 and ctrl[stage].ns == 0
    # Quu <- -1e6 I on one stage of one OCP: G = Quu + B^T P B is negative definite whatever P is
    kkt[sick, stage, L.k_Quu:L.k_Quu + dims.nu ** 2] = (-1.0e6 * np.eye(dims.nu)).reshape(-1)
    rr = RiccatiRecursion(dims, len(ctrl), batch)
    rr.setTimeDiscretization(ctrl)
    rr.backwardRiccatiRecursion(kkt)
    rr.forwardRiccatiRecursion(dx0)
    info = rr.info()
    assert info[sick] & 1, "the failing OCP must carry flag 1"
    assert (np.delete(info, sick) == 0).all(), "the other OCPs must not be flagged"
    with pytest.raises(RuntimeError, match=f"OCP {sick}"):
        rr.checkInfo()
    # the healthy OCPs of the batch are unaffected: they still agree with the oracle
    ok = [b for b in range(batch) if b != sick]
    ric, d = rr.getRiccatiFactorization(), rr.getDirection()
    kk, ric_o, d_o, oinfo = oracle_lib.riccati_batch(dims, L, ctrl, np.ascontiguousarray(kkt[ok]), np.ascontiguousarray(dx0[ok]))
    assert oinfo == 0
    assert rel_err(ric[ok][..., L.r_P:L.r_P + dims.nx ** 2], ric_o[..., L.r_P:L.r_P + dims.nx ** 2]) < 1e-8
    assert rel_err(d[ok][..., L.d_dx:L.d_dx + dims.nx], d_o[..., L.d_dx:L.d_dx + dims.nx]) < 1e-8
    # and a clean sweep afterwards clears the flags
    kkt2, _ = make_kkt(dims, L, ctrl, batch=batch, seed=78)
    rr.backwardRiccatiRecursion(kkt2)
    assert int(rr.info().max()) == 0
    rr.checkInfo()
    rr.close()
