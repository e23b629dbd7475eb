"""GPU parity of the unconstrained stage layer + full UnconstrOCPSolver::updateSolution body vs the CPU oracle
(BASELINE.json configs[0] iiwa14 N=20 batch=1 and configs[1] N=50 batch=256)."""
import numpy as np
import pytest

import oracle_lib
from helpers import rel_err
from robotoc_b200 import UnconstrDirectMultipleShooting, UnconstrRiccatiRecursion, iiwa14_constraint_table
from robotoc_b200.layout import ULayout
from synth import make_unconstr_stage_inputs

pytestmark = pytest.mark.gpu
NV = 7
TOL = 1e-9  # relative, fp64; typical agreement is 1e-14


def _run(N, batch, dt, seed):
    tab = iiwa14_constraint_table()
    rr = UnconstrRiccatiRecursion(NV, N, dt, batch)
    dms = UnconstrDirectMultipleShooting(rr, tab)
    S, UL = dms.layout, rr.layout
    lin, con, sol, dx0 = make_unconstr_stage_inputs(S, N, batch, seed)
    ref = oracle_lib.unconstr_iteration(NV, ULayout(NV, getter=oracle_lib.load().orc_ulayout_get), S, tab, N, dt, lin, con, sol, dx0)
    assert ref["info"] == 0
    dms.condense(lin, con)
    dms.setSolution(sol)
    kkt = dms.getKKT()
    assert rel_err(kkt, ref["kkt"]) < TOL
    assert rel_err(dms.getExpansionData(), ref["ex"]) < TOL
    assert rel_err(dms.getConstraintsData(), ref["con_condensed"]) < TOL
    rr.backwardRiccatiRecursion()
    rr.forwardRiccatiRecursion(dx0)
    d = rr.getDirection()
    assert rel_err(d, ref["dir"]) < 1e-8
    dms.computeStepSizes()
    assert rel_err(dms.getExpandedDirection(), ref["xd"]) < 1e-8
    assert rel_err(dms.getConstraintsData(), ref["con_expanded"]) < 1e-8
    steps = np.stack([dms.maxPrimalStepSize(), dms.maxDualStepSize()], axis=1)
    np.testing.assert_allclose(steps, ref["steps"], rtol=1e-8)
    dms.integrateSolution()
    assert rel_err(dms.getSolution(), ref["sol"]) < 1e-8
    assert rel_err(dms.getConstraintsData(), ref["con"]) < 1e-8
    # the one-call host path gives the same bits as the step-by-step path
    sol2, con2, steps2 = dms.iteration_host(lin, con, sol, dx0)
    np.testing.assert_array_equal(sol2, dms.getSolution())
    np.testing.assert_array_equal(steps2, steps)
    np.testing.assert_array_equal(con2, dms.getConstraintsData())
    assert (con2[:, :N, S.c_slack:S.c_slack + S.nbox] > 0).all()
    rr.close()


def test_unconstr_iteration_iiwa14_n20_batch1():
    _run(20, 1, 0.05, seed=1)


def test_unconstr_iteration_iiwa14_n50_batch256():
    _run(50, 256, 0.02, seed=2)


def test_unconstr_stage_errors():
    rr = UnconstrRiccatiRecursion(NV, 4, 0.1, 2)
    tab = iiwa14_constraint_table()
    bad = iiwa14_constraint_table()
    bad.box[0].idx = 9
    with pytest.raises(ValueError):
        UnconstrDirectMultipleShooting(rr, bad)
    dms = UnconstrDirectMultipleShooting(rr, tab)
    with pytest.raises(ValueError):
        dms.condense(np.zeros((2, 4, dms.layout.l_stride)))  # N+1 = 5 grid points expected
    rr.close()
