"""Every reference solver drives the B200 engine unchanged: the reference's own solver classes (compiled
unmodified into oracle/_ref) run on nbody_engine_b200 (the C++ adapter over the C ABI) and must reproduce
the reference's golden end states -- test/solver/test_nbody_solver.cpp, tolerance 1e-12 absolute."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden_path, load_golden_npz

pytestmark = pytest.mark.gpu

BUTCHER = dict(max_recursion=1, substep_subdivisions=2, refine_steps_count=1, error_threshold=1e-5)
CASES = [
    ("adams5", dict(solver="adams", rank=5)),
    ("adams5-corr", dict(solver="adams", rank=5, correction="true")),
    ("bulirsch-stoer", dict(solver="bs", max_level=4, min_step=1e-5)),
    ("euler", dict(solver="euler")),
    ("midpoint", dict(solver="midpoint")),
    ("midpoint-st", dict(solver="midpoint-st")),
    ("rk4", dict(solver="rk4")),
    ("rkck", dict(solver="rkck", **BUTCHER)),
    ("rkdp", dict(solver="rkdp", correction="false", **BUTCHER)),
    ("rkdp-corr", dict(solver="rkdp", correction="true", **BUTCHER)),
    ("rkdverk", dict(solver="rkdverk", **BUTCHER)),
    ("rkf", dict(solver="rkf", **BUTCHER)),
    ("rkfeagin10", dict(solver="rkfeagin10", **BUTCHER)),
    ("rkfeagin10-corr", dict(solver="rkfeagin10", correction="true", **BUTCHER)),
    ("rkfeagin12", dict(solver="rkfeagin12", **BUTCHER)),
    ("rkfeagin14", dict(solver="rkfeagin14", **BUTCHER)),
    ("rkgl", dict(solver="rkgl", **BUTCHER)),
    ("rklc", dict(solver="rklc", **BUTCHER)),
    ("trapeze2", dict(solver="trapeze", refine_steps_count=2)),
]


@pytest.fixture(scope="module")
def adapter(ref64):
    from nbody_b200 import build
    path = build.adapter_path("f64")
    if not os.path.exists(path):
        pytest.skip("C++ adapter not built (needs the reference headers at build time)")
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    lib.nbody_engine_b200_create.restype = C.c_void_p
    lib.nbody_engine_b200_create.argtypes = [C.c_char_p]
    return lib


def b200_engine(ref64, adapter, **kw):
    from oracle import refharness as R
    h = adapter.nbody_engine_b200_create(R.params(**kw))
    assert h, "factory returned NULL for %r" % (kw,)
    return R.Engine(ref64, handle=h)


def run_golden(ref64, engine, name, params):
    from oracle import refharness as R
    d = R.Data(ref64).load(golden_path("initial_state.txt"))
    assert engine.init(d)
    s = R.Solver(ref64, **params)
    s.set_time_step(1e-3, 3e-2)
    s.set_engine(engine)
    assert s.run(d, 0.3) == 0
    engine.get_data(d)
    expected = R.Data(ref64).load(golden_path(name + ".txt"))
    y, _ = d.export()
    ye, _ = expected.export()
    ok = d.is_equal(expected, 1e-12)
    s.close()          # before the engine: solver destructors free their buffers through engine()
    engine.close()
    return ok, float(np.abs(y - ye).max())


@pytest.mark.parametrize("name,params", CASES, ids=[c[0] for c in CASES])
def test_reference_solver_golden_on_b200(ref64, adapter, name, params):
    ok, err = run_golden(ref64, b200_engine(ref64, adapter, engine="b200"), name, params)
    assert ok, "max |dy| = %g" % err


@pytest.mark.parametrize("kw", [dict(engine="b200", device="0,0"),
                                dict(engine="b200_bh", distance_to_node_radius_ratio=1e8),
                                dict(engine="b200_bh", distance_to_node_radius_ratio=1e8, device="0,0", tree_layout="heap")],
                         ids=["b200-2lanes", "b200_bh-1e8", "b200_bh-1e8-heap-2lanes"])
def test_euler_golden_on_engine_variants(ref64, adapter, kw):
    """The reference re-runs the euler golden on opencl / opencl_bh(1e8) and multi-device lists
    (test_nbody_solver.cpp:267-297); same here for the b200 aliases."""
    ok, err = run_golden(ref64, b200_engine(ref64, adapter, **kw), "euler", dict(solver="euler"))
    assert ok, "max |dy| = %g" % err


def test_adapter_reinit_with_another_body_count_and_early_buffers(ref64, adapter):
    """init() again with a different N (the old state vector is released first), and a buffer created BEFORE the first
    init (ensure_context) that is used as f afterwards."""
    from oracle import refharness as R
    e = b200_engine(ref64, adapter, engine="b200")
    early = e.create_buffer(6 * 128 * 8)
    for stars in (64, 128):
        d = R.Data(ref64).make_universe(stars)
        ref = R.Engine(ref64, engine="simple")
        assert ref.init(d) and e.init(d)
        want = ref.fcompute_y()
        got = e.fcompute_y()
        assert np.abs(got - want).max() <= 1e-13
        if stars == 64:
            e.fcompute(0.0, e.get_y(), early)
            assert np.array_equal(e.read_buffer(early), got)
            e.free_buffer(early)
        ref.close()
        d.close()
    e.close()


def test_factory_rejects_bad_parameters(ref64, adapter):
    from oracle import refharness as R
    for bad in ("", "a", "0,a", "-1", "9999"):
        assert not adapter.nbody_engine_b200_create(R.params(engine="b200", device=bad))
    assert not adapter.nbody_engine_b200_create(R.params(engine="b200_bh", tree_layout="tree"))
    assert not adapter.nbody_engine_b200_create(R.params(engine="cuda"))


def test_adapter_fcompute_vs_simple_engine(ref64, adapter):
    """test_fcompute(e0 = nbody_engine_simple, e1) of test_nbody_engine.cpp:472-558 with the reference's
    default eps 1e-13, N = 128 and 256, two steps with y *= 0.99... (fmadd_inplace(y, y, -0.01) here)."""
    from oracle import refharness as R
    for stars in (64, 128):
        d = R.Data(ref64).make_universe(stars)
        e0 = R.Engine(ref64, engine="simple")
        e1 = b200_engine(ref64, adapter, engine="b200")
        assert e0.init(d) and e1.init(d)
        y0, y1 = e0.create_buffer(e0.size(e0.get_y())), e1.create_buffer(e1.size(e1.get_y()))
        e0.copy_buffer(y0, e0.get_y())
        e1.copy_buffer(y1, e1.get_y())
        nbytes = e0.problem_size() * 8
        for step in range(2):
            f0, f1 = e0.create_buffer(nbytes), e1.create_buffer(nbytes)
            e0.fill_buffer(f0, 1e10)
            e1.fill_buffer(f1, -1e10)
            e0.fcompute(0, y0, f0)
            e1.fcompute(0, y1, f1)
            a, b = e0.read_buffer(f0), e1.read_buffer(f1)
            assert np.abs(a - b).max() <= 1e-13
            e0.free_buffer(f0)
            e1.free_buffer(f1)
            e0.fmadd_inplace(y0, y0, -0.01)
            e1.fmadd_inplace(y1, y1, -0.01)
        assert e1.compute_count() == 2
        for e, y in ((e0, y0), (e1, y1)):
            e.free_buffer(y)
            e.close()
        d.close()


def test_c1_rk4_drift_matches_openmp_engine(ref64, adapter):
    """BASELINE config C1 (--engine=openmp --solver=rk4 --stars_count=1024, G1): the conservation report of
    nbody_data::print_statistics (dP, dL, dE, centre-of-mass drift) after 10 rk4 steps agrees between the
    reference's openmp engine and the b200 engine."""
    from oracle import refharness as R
    stats = {}
    for name, make in (("openmp", lambda: R.Engine(ref64, engine="openmp")),
                       ("b200", lambda: b200_engine(ref64, adapter, engine="b200"))):
        d = R.Data(ref64).make_universe(1024)
        e = make()
        assert e.init(d)
        s = R.Solver(ref64, solver="rk4", max_step=1e-2, min_step=1e-9)
        s.set_engine(e)
        first = d.statistics(e, "PLVE")
        assert s.run(d, 0.1) == 0
        last = d.statistics(e, "PLVE")
        y, _ = d.export()
        stats[name] = (first, last, y, e.compute_count())
        s.close()
        e.close()
        d.close()
    a, b = stats["openmp"], stats["b200"]
    assert a[3] == b[3] and a[3] % 4 == 0 and a[3] >= 40    # 4 fcompute per rk4 step: CC column of the report
    assert np.abs(a[2] - b[2]).max() <= 1e-10               # trajectories agree
    for key in ("E", "P", "L"):
        assert b[1][key] == pytest.approx(a[1][key], rel=1e-11)
    for key in ("dP", "dL", "dE"):
        assert abs(a[1][key] - b[1][key]) <= 1e-9            # per cent of the initial value


def test_adaptive_rkdp_takes_the_same_decisions(ref64, adapter):
    """Error-controlled rkdp (fmaxabs of the embedded error -> host branch -> step subdivision,
    nbody_solver_rk_butcher.cpp:213-231) on b200 vs the reference's openmp engine: the same number of fcompute calls
    (i.e. the same accept/subdivide decisions) and the same trajectory."""
    from oracle import refharness as R
    out = {}
    for name, make in (("openmp", lambda: R.Engine(ref64, engine="openmp")),
                       ("b200", lambda: b200_engine(ref64, adapter, engine="b200")),
                       ("b200-2lanes", lambda: b200_engine(ref64, adapter, engine="b200", device="0,0"))):
        d = R.Data(ref64).make_universe(512)                 # N = 1024
        e = make()
        assert e.init(d)
        s = R.Solver(ref64, solver="rkdp", max_step=5e-2, min_step=1e-9, error_threshold=1e-6, max_recursion=6,
                     substep_subdivisions=2, refine_steps_count=1)
        s.set_engine(e)
        assert s.run(d, 0.25) == 0
        e.get_data(d)
        y, _ = d.export()
        out[name] = (e.compute_count(), y)
        s.close()
        e.close()
        d.close()
    assert out["openmp"][0] == out["b200"][0] == out["b200-2lanes"][0]
    assert out["openmp"][0] > 7 * 5                          # subdivision really happened
    assert np.abs(out["openmp"][1] - out["b200"][1]).max() <= 1e-9
    assert np.array_equal(out["b200"][1], out["b200-2lanes"][1]) or np.abs(out["b200"][1] - out["b200-2lanes"][1]).max() <= 1e-12


def test_fp32_adapter_euler_vs_reference_fp32(ref32):
    """NB_COORD_PRECISION=1 twin: libnbody_engine_b200_f32 + libnb200_f32 driven by the FP32 build of the reference's
    euler solver; trajectory vs the reference's own FP32 simple engine."""
    from nbody_b200 import build
    from oracle import refharness as R
    path = build.adapter_path("f32")
    if not os.path.exists(path):
        pytest.skip("FP32 adapter not built")
    ad = C.CDLL(path, mode=C.RTLD_LOCAL)
    ad.nbody_engine_b200_create.restype = C.c_void_p
    ad.nbody_engine_b200_create.argtypes = [C.c_char_p]
    out = {}
    for name in ("simple", "b200", "b200_bh"):
        d = R.Data(ref32).make_universe(128)
        if name == "simple":
            e = R.Engine(ref32, engine="simple")
        else:
            kw = dict(engine=name) if name == "b200" else dict(engine=name, distance_to_node_radius_ratio=1e8)
            h = ad.nbody_engine_b200_create(R.params(**kw))
            assert h
            e = R.Engine(ref32, handle=h)
        assert e.init(d)
        s = R.Solver(ref32, solver="euler", max_step=1e-2)
        s.set_engine(e)
        assert s.run(d, 0.05) == 0
        e.get_data(d)
        y, _ = d.export()
        out[name] = y.astype(np.float64)
        s.close()
        e.close()
        d.close()
    scale = np.abs(out["simple"]).max()
    assert np.abs(out["b200"] - out["simple"]).max() <= 2e-5 * scale
    assert np.abs(out["b200_bh"] - out["simple"]).max() <= 2e-5 * scale


@pytest.mark.parametrize("kind,path", [("G1", "initial_state.txt"), ("SI", "initial_state.txt"), ("ADK", "initial_state.txt"),
                                       ("Zeno", "zeno_ascii.txt")])
def test_initial_state_types_direct_and_barnes_hut(ref64, adapter, kind, path):
    """--initial_type G1 / SI / ADK / Zeno (nbody_data::load_initial, nbody_data.cpp:618-660: the unit system scales
    the masses by G): the same loaded nbody_data on the reference's CPU engines and on the b200 aliases -- per-body
    acceleration <= 1e-12 relative (direct vs openmp, Barnes-Hut vs simple_bh with the same tree layout and
    opening ratio), and a short rk4 run ends in the same state."""
    from oracle import refharness as R
    bh = dict(distance_to_node_radius_ratio=10, tree_layout="heap_stackless")
    # (the reference's block engine needs N % 64 == 0 -- these files hold 16 bodies -- so openmp is the direct oracle here)
    pairs = [(dict(engine="openmp"), dict(engine="b200")),
             (dict(engine="simple_bh", traverse_type="nested_tree", **bh), dict(engine="b200_bh", **bh))]
    for ref_kw, our_kw in pairs:
        out = []
        for make in (lambda: R.Engine(ref64, **ref_kw), lambda: b200_engine(ref64, adapter, **our_kw)):
            d = R.Data(ref64).load_initial(golden_path(path), kind)
            n = d.count
            e = make()
            assert e.init(d)
            f = e.create_buffer(e.size(e.get_y()))
            e.fcompute(0, e.get_y(), f)
            acc = e.read_buffer(f).reshape(6, n)[3:]
            e.free_buffer(f)
            s = R.Solver(ref64, solver="rk4", max_step=1e-3, min_step=1e-9)
            s.set_engine(e)
            assert s.run(d, 5e-3) == 0
            e.get_data(d)
            out.append((acc, d.export()[0].copy()))
            s.close()
            e.close()
            d.close()
        (a_ref, y_ref), (a_our, y_our) = out
        rel = np.sqrt(((a_ref - a_our) ** 2).sum(0)) / np.sqrt((a_ref ** 2).sum(0))
        assert rel.max() <= 1e-12, "%s %r: %g" % (kind, our_kw, rel.max())
        assert np.abs(y_ref - y_our).max() <= 1e-12 * max(1.0, np.abs(y_ref).max())
