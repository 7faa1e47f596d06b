"""The reference's OWN factory, patched with integration/nbody_engines.patch, creates the nb200 engines (run under the
stand-in CUDA runtime by tests/test_integration_patch.py:  drive_factory.py <mock_cudart.so> <libfactory.so>)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")
from oracle import refharness as R  # noqa: E402


def main():
    lib = R.load("f64")
    fac = C.CDLL(sys.argv[2])
    fac.factory_create.restype = C.c_void_p
    fac.factory_create.argtypes = [C.c_char_p]
    want = {b"engine=b200": "nbody_engine_b200", b"engine=b200;device=0,0": "nbody_engine_b200",
            b"engine=b200_bh;distance_to_node_radius_ratio=4;tree_layout=heap": "nbody_engine_b200_bh",
            b"engine=openmp": "nbody_engine_openmp", b"engine=block": "nbody_engine_block",
            b"engine=simple_bh;traverse_type=nested_tree;tree_layout=heap_stackless": None}
    for params, name in want.items():
        h = fac.factory_create(params)
        assert h, params
        e = R.Engine(lib, handle=h)
        if name is not None:
            assert e.type_name() == name, (params, e.type_name())
        e.close()
    for bad in (b"engine=b200;device=", b"engine=b200;device=a", b"engine=b200;device=99", b"engine=b200_bh;tree_layout=tree",
                b"engine=no_such_engine"):
        assert not fac.factory_create(bad), bad
    # and the engine the patched factory made drives a reference solver
    d = R.Data(lib).load(os.path.join(ROOT, "tests", "golden", "initial_state.txt"))
    e = R.Engine(lib, handle=fac.factory_create(b"engine=b200"))
    assert e.init(d)
    s = R.Solver(lib, solver="rk4")
    s.set_time_step(1e-3, 3e-2)
    s.set_engine(e)
    assert s.run(d, 0.3) == 0
    assert e.compute_count() == 40
    s.close()
    e.close()
    d.close()
    print("patched factory ok")


if __name__ == "__main__":
    main()
