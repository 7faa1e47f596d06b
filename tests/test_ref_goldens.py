"""Pins oracle/_ref (the reference compiled unmodified against the Qt shim): its own solvers on its own
nbody_engine_simple reproduce the 19 golden end states of test/data within the reference's 1e-12 gate
(test/solver/test_nbody_solver.cpp:43-85). Runs on CPU; skipped where oracle/_ref was not built."""
import numpy as np
import pytest

from conftest import golden_path
from test_solvers_gpu import CASES


@pytest.mark.parametrize("name,params", CASES, ids=[c[0] for c in CASES])
def test_compiled_reference_reproduces_golden(ref64, name, params):
    from oracle import refharness as R
    d = R.Data(ref64).load(golden_path("initial_state.txt"))
    e = R.Engine(ref64, engine="simple")
    assert e.init(d)
    s = R.Solver(ref64, **params)
    s.set_time_step(1e-3, 3e-2)
    s.set_engine(e)
    assert s.run(d, 0.3) == 0
    e.get_data(d)
    expected = R.Data(ref64).load(golden_path(name + ".txt"))
    assert d.is_equal(expected, 1e-12)
    chk = s.butcher_check()
    if chk is not None:            # butcher_table_check (:87-144): sum b = 1, sum_j a_ij = c_i
        eps = 10 * np.finfo(np.float64).eps
        assert abs(chk[0] - 1) < eps and abs(chk[1] - 1) < eps and chk[2] < eps
    s.close()
    e.close()
    d.close()
    expected.close()


def test_invalid_solver_names_are_rejected(ref64):
    from oracle import refharness as R
    with pytest.raises(ValueError):
        R.Solver(ref64, solver="invalid")
    with pytest.raises(ValueError):
        R.Solver(ref64, solver="adams", starter_solver="invalid")


def test_reference_cross_engine_gates(ref64):
    """Cross-engine thresholds of test_nbody_engine.cpp:1358-1558 hold for the compiled reference:
    heap == heap_stackless (1e-16), BH(1e8) vs direct 1e-11."""
    from oracle import refharness as R
    d = R.Data(ref64).make_universe(128)
    out = {}
    for key, kw in (("simple", dict(engine="simple")), ("block", dict(engine="block")),
                    ("heap", dict(engine="simple_bh", tree_layout="heap", traverse_type="nested_tree", distance_to_node_radius_ratio=3.1623)),
                    ("stackless", dict(engine="simple_bh", tree_layout="heap_stackless", traverse_type="nested_tree", distance_to_node_radius_ratio=3.1623)),
                    ("bh1e8", dict(engine="simple_bh", tree_layout="heap_stackless", traverse_type="nested_tree", distance_to_node_radius_ratio=1e8))):
        e = R.Engine(ref64, **kw)
        assert e.init(d)
        out[key] = e.fcompute_y()
        e.close()
    assert np.abs(out["heap"] - out["stackless"]).max() <= 1e-16
    assert np.abs(out["bh1e8"] - out["simple"]).max() <= 1e-11
    assert np.abs(out["block"] - out["simple"]).max() <= 1e-11
    d.close()
