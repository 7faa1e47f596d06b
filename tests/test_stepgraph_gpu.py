"""SURVEY 8f ranks 2 and 3 on the GPU, through the C ABI:

* solver steps as CUDA graphs (``step_graph``): a fixed-step solver replayed as one graph launch per step gives
  BIT-IDENTICAL states to the eager engine; adaptive solvers (fmaxabs inside the step), changing coefficients,
  rotating buffers and mid-step reads fall back to eager issue without changing a bit;
* body arrays <-> state vector (``nb200_write_bodies`` / ``nb200_read_bodies``): the AoS transposes of
  nbody_engine_cuda::init / get_data (nbody_engine_cuda.cpp:113-175) done on the device.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden_path, load_golden_npz
from test_solvers_gpu import BUTCHER, CASES, adapter, b200_engine, run_golden  # noqa: F401  (adapter is a fixture)

pytestmark = pytest.mark.gpu


# ---- a fixed-step solver written against the engine API (the shape of nbody_solver_rk4.cpp:30-62) ------------------
class Rk4:
    def __init__(self, e):
        self.e = e
        size = e.get_y().size()
        self.k = e.create_buffers(size, 4)
        self.tmp = e.create_buffer(size)

    def advise(self, dt):
        e, k, y = self.e, self.k, self.e.get_y()
        t = e.get_time()
        e.fcompute(t, y, k[0])
        e.fmadd(self.tmp, y, k[0], 0.5 * dt)
        e.fcompute(t + 0.5 * dt, self.tmp, k[1])
        e.fmadd(self.tmp, y, k[1], 0.5 * dt)
        e.fcompute(t + 0.5 * dt, self.tmp, k[2])
        e.fmadd(self.tmp, y, k[2], dt)
        e.fcompute(t + dt, self.tmp, k[3])
        coeff = np.array([dt / 6, dt / 3, dt / 3, dt / 6], dtype=e.dtype)
        e.fmaddn_inplace(y, k, coeff)
        e.advise_time(dt)

    def close(self):
        self.e.free_buffers(self.k)
        self.e.free_buffer(self.tmp)


def run_rk4(precision, kind, steps, graph, dts=None, peek_at=None, n_tag="g1_n2048", **kw):
    from nbody_b200 import Engine
    g = load_golden_npz(n_tag, precision)
    with Engine(precision=precision, devices="0", kind=kind, **kw) as e:
        assert e.init(g["y"], g["mass"])
        if graph:
            assert e.set_option("step_graph", 1) == 0
        s = Rk4(e)
        peeked = []
        for i in range(steps):
            dt = 1e-3 if dts is None else dts[i]
            if peek_at is not None and i == peek_at:
                # a host-visible read in the MIDDLE of a step: everything issued before it must have run
                y, t = e.get_y(), e.get_time()
                e.fcompute(t, y, s.k[0])
                peeked.append(e.read_buffer(s.k[0]))
                e.fmadd(s.tmp, y, s.k[0], 0.5 * dt)
                e.fcompute(t, s.tmp, s.k[1])
                e.fmadd(s.tmp, y, s.k[1], 0.5 * dt)
                e.fcompute(t, s.tmp, s.k[2])
                e.fmadd(s.tmp, y, s.k[2], dt)
                e.fcompute(t, s.tmp, s.k[3])
                e.fmaddn_inplace(y, s.k, np.array([dt / 6, dt / 3, dt / 3, dt / 6], dtype=e.dtype))
                e.advise_time(dt)
            else:
                s.advise(dt)
        out = e.read_buffer(e.get_y())
        stats = e.step_graph_stats()
        launches = e.launch_count()
        s.close()
    return out, stats, launches, peeked


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("kind,kw", [("direct", {}), ("bh", dict(distance_to_node_radius_ratio=10.0)),
                                     ("bh", dict(distance_to_node_radius_ratio=10.0, tree_build_rate=4))],
                         ids=["direct", "bh", "bh-build-rate-4"])
def test_rk4_replayed_as_graphs_is_bit_identical(precision, kind, kw):
    steps = 12
    eager, st0, launches0, _ = run_rk4(precision, kind, steps, graph=False, **kw)
    graph, st1, launches1, _ = run_rk4(precision, kind, steps, graph=True, **kw)
    assert np.array_equal(eager, graph)
    assert st0["state"] == "off" and st0["graph_launches"] == 0
    if kw.get("tree_build_rate"):
        # every 4th step rebuilds the tree, the others refresh it: two distinct steps, both end up with a graph, and a
        # wrong guess between them is corrected at the first call (no replay is abandoned)
        assert st1["distinct_steps"] == 2 and st1["bailouts"] == 0
        assert st1["graph_launches"] >= steps - 4
    else:
        assert st1["state"] == "replay" and st1["distinct_steps"] == 1
        assert st1["graph_launches"] == steps - 2          # two steps recorded eagerly (the second confirms the first
        assert st1["bailouts"] == 0                          # repeats), the third captured, the rest replayed
        assert st1["launches_per_step"] >= 8                  # direct at N = 2,048: 4 single-launch fcomputes + 4 state ops
        assert launches1 == launches0                        # the same kernels ran, only issued differently


def test_changed_step_size_falls_back_and_recaptures():
    steps = 12
    dts = [1e-3] * 5 + [5e-4] * 7
    eager, _, _, _ = run_rk4("f64", "direct", steps, graph=False, dts=dts)
    graph, st, _, _ = run_rk4("f64", "direct", steps, graph=True, dts=dts)
    assert np.array_equal(eager, graph)
    assert st["bailouts"] == 1                               # step 6: the first fmadd carries another coefficient
    assert st["state"] == "replay" and st["graph_launches"] >= 7 and st["distinct_steps"] == 2


def test_read_in_the_middle_of_a_replayed_step():
    steps = 8
    eager, _, _, peek0 = run_rk4("f64", "direct", steps, graph=False, peek_at=5)
    graph, st, _, peek1 = run_rk4("f64", "direct", steps, graph=True, peek_at=5)
    assert np.array_equal(eager, graph)
    assert len(peek0) == 1 and np.array_equal(peek0[0], peek1[0])
    assert st["bailouts"] == 1


def test_every_step_different_is_never_captured():
    steps = 14
    dts = [1e-3 * (1 + 0.01 * i) for i in range(steps)]
    eager, _, launches0, _ = run_rk4("f64", "direct", steps, graph=False, dts=dts)
    graph, st, launches1, _ = run_rk4("f64", "direct", steps, graph=True, dts=dts)
    assert np.array_equal(eager, graph)
    # nothing ever repeats, so nothing is predicted: every step runs eagerly, no capture is wasted
    assert st["state"] == "record" and st["graph_launches"] == 0 and st["bailouts"] == 0
    assert st["distinct_steps"] == steps and launches0 == launches1


def test_alternating_step_sizes_get_one_graph_each():
    """A caller that alternates between two steps (here: two step sizes; Bulirsch-Stoer's sub-steps and Adams' rotating
    history buffers are the reference's cases): both are filed, both get a graph, and from then on every step is one
    graph launch."""
    steps = 16
    dts = [1e-3 if i % 2 == 0 else 5e-4 for i in range(steps)]
    eager, _, _, _ = run_rk4("f64", "direct", steps, graph=False, dts=dts)
    graph, st, _, _ = run_rk4("f64", "direct", steps, graph=True, dts=dts)
    assert np.array_equal(eager, graph)
    assert st["distinct_steps"] == 2 and st["state"] == "replay"
    assert st["graph_launches"] >= steps - 6 and st["bailouts"] <= 1


class Rk4Err(Rk4):
    """rk4 with an error estimate read back in the middle of every step, the call pattern of the reference's embedded
    Runge-Kutta solvers (nbody_solver_rk_butcher.cpp:207-231): fmaddn -> fmaxabs -> host branch -> commit or subdivide."""

    def __init__(self, e, threshold):
        super().__init__(e)
        self.err = e.create_buffer(e.get_y().size())
        self.threshold = threshold
        self.errors = []
        self.halved = 0

    def advise(self, dt):
        e, k, y = self.e, self.k, self.e.get_y()
        t = e.get_time()
        e.fcompute(t, y, k[0])
        e.fmadd(self.tmp, y, k[0], 0.5 * dt)
        e.fcompute(t + 0.5 * dt, self.tmp, k[1])
        e.fmadd(self.tmp, y, k[1], 0.5 * dt)
        e.fcompute(t + 0.5 * dt, self.tmp, k[2])
        e.fmadd(self.tmp, y, k[2], dt)
        e.fcompute(t + dt, self.tmp, k[3])
        e.fmaddn(self.err, None, k, np.array([dt / 6, -dt / 6, -dt / 6, dt / 6], dtype=e.dtype))
        err = e.fmaxabs(self.err)
        self.errors.append(err)
        if err > self.threshold:
            # "subdivide": another call sequence after the border
            self.halved += 1
            e.copy_buffer(self.tmp, y)
            e.fmadd_inplace(self.tmp, k[0], 0.5 * dt)
            e.fcompute(t + 0.5 * dt, self.tmp, k[1])
            e.fmadd_inplace(y, k[1], dt)
        else:
            e.fmaddn_inplace(y, k, np.array([dt / 6, dt / 3, dt / 3, dt / 6], dtype=e.dtype))
        e.advise_time(dt)

    def close(self):
        self.e.free_buffer(self.err)
        super().close()


def run_rk4_err(graph, steps, threshold):
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n2048", "f64")
    with Engine(precision="f64", devices="0") as e:
        assert e.init(g["y"], g["mass"])
        if graph:
            assert e.set_option("step_graph", 1) == 0
        s = Rk4Err(e, threshold)
        for _ in range(steps):
            s.advise(1e-3)
        out = e.read_buffer(e.get_y())
        stats, launches = e.step_graph_stats(), e.launch_count()
        errors, halved = list(s.errors), s.halved
        s.close()
    return out, stats, launches, errors, halved


def test_fmaxabs_inside_the_step_splits_the_graph_in_two():
    steps = 10
    eager, _, launches0, err0, _ = run_rk4_err(False, steps, threshold=1e30)
    graph, st, launches1, err1, _ = run_rk4_err(True, steps, threshold=1e30)
    assert np.array_equal(eager, graph)
    assert err0 == err1 and all(v > 0 for v in err0)         # the reduction saw the finished stages every time
    assert st["state"] == "replay" and st["bailouts"] == 0
    assert st["graph_launches"] == 2 * (steps - 2)           # two segments per step: before and after the border
    assert launches0 == launches1


def test_solver_branching_after_fmaxabs_leaves_the_replay():
    """The recorded step commits; when the error estimate first exceeds the threshold the solver issues other calls
    after the border: the first segment has run as a graph, the rest of that step runs eagerly."""
    steps = 12
    _, _, _, errors, _ = run_rk4_err(False, steps, threshold=1e30)
    threshold = sorted(errors)[steps // 2]                   # about half of the steps "subdivide"
    eager, _, _, err0, halved0 = run_rk4_err(False, steps, threshold)
    graph, st, _, err1, halved1 = run_rk4_err(True, steps, threshold)
    assert 0 < halved0 < steps and halved0 == halved1
    assert np.array_equal(eager, graph)
    assert err0 == err1


def test_step_graph_is_ignored_with_lanes():
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n128", "f64")
    with Engine(precision="f64", devices="0,0") as e:
        assert e.init(g["y"], g["mass"])
        assert e.set_option("step_graph", 1) == 0
        s = Rk4(e)
        for _ in range(4):
            s.advise(1e-3)
        assert e.step_graph_stats()["state"] == "off"
        s.close()


# ---- the reference's own solvers, unchanged, on the adapter with step_graph=1 ---------------------------------------
@pytest.mark.parametrize("name,params", CASES, ids=[c[0] for c in CASES])
def test_reference_solver_golden_with_step_graphs(ref64, adapter, name, params):
    """test/solver/test_nbody_solver.cpp goldens, 1e-12 absolute, with every solver step offered to the graph path;
    and the end state equals the eager engine's bit for bit."""
    from oracle import refharness as R
    states = {}
    for graph in (0, 1):
        e = b200_engine(ref64, adapter, engine="b200", step_graph=graph)
        d = R.Data(ref64).load(golden_path("initial_state.txt"))
        assert e.init(d)
        s = R.Solver(ref64, **params)
        s.set_time_step(1e-3, 3e-2)
        s.set_engine(e)
        assert s.run(d, 0.3) == 0
        e.get_data(d)
        out = (C.c_ulonglong * 5)()
        adapter.nbody_engine_b200_step_graph_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        assert adapter.nbody_engine_b200_step_graph_stats(e.h, out) == 0
        expected = R.Data(ref64).load(golden_path(name + ".txt"))
        assert d.is_equal(expected, 1e-12)
        states[graph] = (d.export()[0].copy(), list(out))
        s.close()
        e.close()
    assert np.array_equal(states[0][0], states[1][0])
    assert states[0][1][0] == 0
    if name in ("euler", "rk4"):
        assert states[1][1][0] >= 5, "fixed-step solver was not replayed: %r" % (states[1][1],)
    if name in ("rkf", "rkdp", "rkfeagin14"):
        # embedded tables read the error norm back in every step: replayed as two graphs per step around the border
        assert states[1][1][0] >= 4, "embedded solver was not replayed: %r" % (states[1][1],)
    if name in ("adams5", "bulirsch-stoer"):
        # periodic patterns (rotating history buffers; sub-steps of several sizes): several distinct steps, some replayed
        assert states[1][1][4] >= 2


# ---- body arrays <-> state vector -----------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,devices", [("f64", "0"), ("f32", "0"), ("f64", "0,0"), ("f64", "0,0,0,0")],
                         ids=["f64", "f32", "f64-2lanes", "f64-4lanes"])
@pytest.mark.parametrize("n", [4, 1000, 2048, 65536])
def test_bodies_round_trip(precision, devices, n):
    from nbody_b200 import Engine
    dtype = np.float64 if precision == "f64" else np.float32
    rng = np.random.RandomState(n)
    pos = rng.standard_normal((n, 3)).astype(dtype)
    vel = rng.standard_normal((n, 3)).astype(dtype)
    mass = rng.uniform(0.5, 1.5, n).astype(dtype)
    with Engine(precision=precision, devices=devices) as e:
        assert e.init_bodies(pos, vel, mass)
        y = e.read_buffer(e.get_y())
        want = np.concatenate([pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2]])
        assert np.array_equal(y, want)                       # write_bodies == the reference's host transpose
        p2 = np.full((n, 3), np.nan, dtype=dtype)
        v2 = np.full((n, 3), np.nan, dtype=dtype)
        assert e.host_register(p2) == 0 and e.host_register(v2) == 0
        assert e.get_bodies(p2, v2) is not None
        assert np.array_equal(p2, pos) and np.array_equal(v2, vel)
        e.host_unregister(p2)
        e.host_unregister(v2)
        # any state-sized buffer can be read as bodies (get_data after a solver step reads m_y)
        f = e.create_buffer(e.get_y().size())
        e.fmadd(f, e.get_y(), e.get_y(), 1.0)
        p3, v3 = e.get_bodies(y=f)
        assert np.array_equal(p3, 2 * pos) and np.array_equal(v3, 2 * vel)
        # negative branches: log and return
        small = e.create_buffer(64)
        assert e.get_bodies(y=small) is None
        assert e.get_bodies(np.zeros((n, 2), dtype=dtype), v2) is None
        e.free_buffer(small)
        e.free_buffer(f)


def test_barnes_hut_with_launch_order_inside_the_graph():
    """N = 262,144 Barnes-Hut euler steps: the walk's longest-walk-first launch order (costs written by the walk, sorted on
    the device, read by the next walk) lives inside the replayed graph. Bit-identical to eager issue with natural order."""
    from nbody_b200 import Engine
    from util import universe
    y, m = universe(262144)
    out = {}
    for graph, lpt in ((0, 0), (1, -1)):
        with Engine(kind="bh", distance_to_node_radius_ratio=2.0) as e:
            e.set_option("walk_lpt", lpt)
            assert e.init(y, m)
            e.set_option("step_graph", graph)
            dy = e.create_buffer(e.get_y().size())
            for _ in range(6):
                e.fcompute(e.get_time(), e.get_y(), dy)
                e.fmadd_inplace(e.get_y(), dy, 1e-3)
                e.advise_time(1e-3)
            out[graph] = (e.read_buffer(e.get_y()), e.step_graph_stats())
    assert np.array_equal(out[0][0], out[1][0])
    assert out[1][1]["state"] == "replay" and out[1][1]["graph_launches"] == 4 and out[1][1]["bailouts"] == 0
