"""Generates tests/golden/golden_ref_stage_r2.npz: one FULL hot-path iteration computed by the reference's own code
(tests/ref_lib.reference_iteration: /root/reference/src/dynamics, src/constraints, src/core and src/riccati compiled unmodified,
oracle/Makefile.ref) on the small event schedule with switching-time optimisation (Intermediate, Lift, switching-constraint and
Impact stages), batch 1.  Only runs where /root/reference exists:   python tests/golden/make_golden_ref_stage.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import small_event_schedule  # noqa: E402
from robotoc_b200 import ANYMAL  # noqa: E402
from robotoc_b200.stage import StageDims, anymal_constraint_table  # noqa: E402
from synth import make_stage_inputs  # noqa: E402

SEED, BATCH = 301, 1
KEYS = ("kkt", "cc_cond", "ric", "d", "cc_exp", "xd_exp", "steps", "d_upd", "xd_upd", "cc_upd", "ex_upd", "perf_stage")


def problem(S_getter=None, K_getter=None, impact_cones=False):
    """impact_cones: the same problem with ImpactFrictionCone registered (examples/anymal/run.cpp:173-181): the impact stage
    carries friction-cone rows on the impact forces (golden keys prefixed "ic_")."""
    from robotoc_b200 import Layout, StageLayout
    table = anymal_constraint_table(impact_friction_cone=impact_cones)
    sd = StageDims(ANYMAL, nf_max=12, n_contacts=table.n_contacts, n_box=table.n_box)
    S, K = StageLayout(sd, getter=S_getter), Layout(ANYMAL, getter=K_getter)
    td, ev, ctrl = small_event_schedule(True)
    lin, con, sol, dx0 = make_stage_inputs(sd, S, ctrl, BATCH, SEED, impact_cones=impact_cones)
    return table, sd, S, K, ctrl, lin, con, sol, dx0


if __name__ == "__main__":
    import oracle_lib
    import ref_lib
    lib = oracle_lib.load()
    table, sd, S, K, ctrl, lin, con, sol, dx0 = problem(lib.orc_stage_layout_get, lib.orc_layout_get)
    out = ref_lib.reference_iteration(sd, S, K, table, ctrl, lin, con, dx0)
    data = {k: out[k] for k in KEYS}
    table, sd, S, K, ctrl, lin, con, sol, dx0 = problem(lib.orc_stage_layout_get, lib.orc_layout_get, impact_cones=True)
    out = ref_lib.reference_iteration(sd, S, K, table, ctrl, lin, con, dx0)
    data.update({"ic_" + k: out[k] for k in KEYS})
    path = os.path.join(HERE, "golden_ref_stage_r2.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
