"""Host-logic scenarios for libnb200 on a machine WITHOUT a GPU (test infrastructure; run by tests/test_host_mock_cuda.py
as  LD_PRELOAD=<mock_cudart.so> python tests/mock_cuda/drive.py <mock_cudart.so>).

Under the mock runtime "device" memory is host memory, copies are real, kernels are only counted and a captured graph
replays its copies and counts its kernels (tests/mock_cuda/mock_cudart.c). That is enough to pin down what no CPU oracle
can: which calls the library issues WHEN -- the shard layout behind write/read_buffer, the argument checks, and the step
table of nb200_stepgraph.cuh (eager / captured / replayed launches per solver step, deferral, fall-backs)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nbody_b200 import Engine  # noqa: E402

MOCK = C.CDLL(sys.argv[1])
EAGER, CAPTURED, GRAPHS, REPLAYED, INSTANTIATED, COPIES, CAPTURES, ALLOCS = range(8)


def counters():
    out = (C.c_ulonglong * 8)()
    MOCK.mock_counters(out)
    return np.array(list(out), dtype=np.int64)


class Delta:
    """Counter increments over a with-block."""

    def __enter__(self):
        self.before = counters()
        return self

    def __exit__(self, *exc):
        self.d = counters() - self.before


def system(n=64):
    rng = np.random.RandomState(n)
    return rng.uniform(-1, 1, 6 * n), rng.uniform(0.5, 1.5, n)


class Rk4:
    """The call pattern of nbody_solver_rk4.cpp:30-62 (N <= 4096 on one shard: one kernel per call, 8 per step)."""

    def __init__(self, e, with_error_norm=False):
        self.e, self.norm = e, with_error_norm
        size = e.get_y().size()
        self.k = e.create_buffers(size, 4)
        self.tmp = e.create_buffer(size)

    def advise(self, dt, peek=False):
        e, k, y = self.e, self.k, self.e.get_y()
        e.fcompute(0.0, y, k[0])
        if peek:
            e.read_buffer(k[0])
        e.fmadd(self.tmp, y, k[0], 0.5 * dt)
        e.fcompute(0.0, self.tmp, k[1])
        e.fmadd(self.tmp, y, k[1], 0.5 * dt)
        e.fcompute(0.0, self.tmp, k[2])
        e.fmadd(self.tmp, y, k[2], dt)
        e.fcompute(0.0, self.tmp, k[3])
        if self.norm:
            e.fmaddn(self.tmp, None, k, np.array([dt, -dt, -dt, dt]))
            e.fmaxabs(self.tmp)
        e.fmaddn_inplace(y, k, np.array([dt / 6, dt / 3, dt / 3, dt / 6]))
        e.advise_time(dt)


def scenario_buffers_and_shards():
    y, m = system(64)
    for devices in ("0", "0,1", "0,0,0,0"):
        with Engine(devices=devices) as e:
            assert e.init(y, m)
            assert np.array_equal(e.read_buffer(e.get_y()), y)                 # 6 rows x N -> shards of 6 rows x N/G and back
            a, b = e.create_buffer(y.nbytes), e.create_buffer(y.nbytes)
            e.write_buffer(a, 2 * y)
            e.copy_buffer(b, a)
            assert np.array_equal(e.read_buffer(b), 2 * y)
            odd = e.create_buffer(33 * 8)                                       # not a state vector: replicated on every lane
            data = np.arange(33.0)
            e.write_buffer(odd, data)
            assert np.array_equal(e.read_buffer(odd), data)
            lib, ctx = e.lib, e.ctx
            assert lib.nb200_copy(ctx, odd.handle, a.handle) == -1             # NB200_ERR_ARG: sizes differ
            assert lib.nb200_fmadd_inplace(ctx, a.handle, odd.handle, 1.0) == -1
            assert lib.nb200_fcompute_direct(ctx, a.handle, a.handle) == -1    # y and f must differ
            assert lib.nb200_fcompute_direct(ctx, odd.handle, a.handle) == -1
            assert lib.nb200_free(ctx, C.c_void_p(12345)) == -1                # foreign handle
            assert lib.nb200_free(ctx, None) == 0                              # free_buffer(nullptr) is a no-op
            with Engine(devices="0") as other:
                assert other.init(y, m)
                assert lib.nb200_copy(ctx, a.handle, other.get_y().handle) == -1   # a buffer of another context
            before = counters()[ALLOCS]
            e.free_buffer(a), e.free_buffer(b), e.free_buffer(odd)
            assert counters()[ALLOCS] == before - 3 * len(devices.split(","))
    print("ok buffers_and_shards")


def scenario_fixed_step_replay():
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        s = Rk4(e)
        per_step = []
        for i in range(8):
            before = e.launch_count()
            with Delta() as c:
                s.advise(1e-3)
            per_step.append((c.d[EAGER], c.d[CAPTURED], c.d[GRAPHS], c.d[REPLAYED], c.d[CAPTURES]))
            assert e.launch_count() - before == 8                              # the library's own launch counter never drifts
        assert per_step[0] == per_step[1] == (8, 0, 0, 0, 0)                   # recorded; the second step confirms the first
        assert per_step[2] == (0, 8, 1, 8, 1)                                  # captured, and run by launching the capture
        assert all(p == (0, 0, 1, 8, 0) for p in per_step[3:])                 # one graph launch per step, nothing else
        st = e.step_graph_stats()
        assert st["state"] == "replay" and st["distinct_steps"] == 1 and st["bailouts"] == 0 and st["launches_per_step"] == 8
    print("ok fixed_step_replay")


def scenario_deferral_is_real():
    """copy_buffer inside a replayed step runs when the graph is launched (the boundary), not when it is called; a read in
    between forces everything accepted so far to run first."""
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        src, dst = e.create_buffer(y.nbytes), e.create_buffer(y.nbytes)

        def step(value):
            e.write_buffer(src, np.full(y.size, float(value)))                 # between steps: harmless
            e.copy_buffer(dst, src)
            e.fmadd_inplace(src, dst, 0.0)
            e.advise_time(1e-3)

        for v in range(4):
            step(v)
        assert e.step_graph_stats()["state"] == "replay"
        e.write_buffer(src, np.full(y.size, 7.0))
        with Delta() as c:
            e.copy_buffer(dst, src)                                            # accepted and deferred: nothing is copied yet
        assert c.d[COPIES] == 0 and c.d[EAGER] == 0
        with Delta() as c:
            host = e.read_buffer(dst)                                          # a caller's only way to look: it flushes first
        assert np.all(host == 7.0) and c.d[GRAPHS] == 0 and c.d[COPIES] == 2   # the copy, issued eagerly, then the read
        e.fmadd_inplace(src, dst, 0.0)
        e.advise_time(1e-3)
        assert e.step_graph_stats()["bailouts"] == 1
        for v in range(10, 14):
            step(v)
        assert e.step_graph_stats()["state"] == "replay"
        with Delta() as c:
            step(99)
        assert c.d[GRAPHS] == 1 and c.d[EAGER] == 0 and np.all(e.read_buffer(dst) == 99.0)
    print("ok deferral_is_real")


def scenario_error_norm_border():
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        s = Rk4(e, with_error_norm=True)
        for _ in range(3):
            s.advise(1e-3)
        with Delta() as c:
            s.advise(1e-3)
        # two segments around fmaxabs; the reduction itself (one kernel) runs when called
        assert c.d[GRAPHS] == 2 and c.d[EAGER] == 1 and c.d[REPLAYED] == 9 and c.d[CAPTURES] == 0
        assert e.step_graph_stats()["bailouts"] == 0
    print("ok error_norm_border")


def scenario_read_in_the_middle():
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        s = Rk4(e)
        for _ in range(5):
            s.advise(1e-3)
        with Delta() as c:
            s.advise(1e-3, peek=True)
        assert c.d[EAGER] == 8 and c.d[GRAPHS] == 0                            # 1 accepted call issued at the read + 7 eager
        assert e.step_graph_stats()["bailouts"] == 1
        s.advise(1e-3)                                                          # the chain was broken: recorded again
        with Delta() as c:
            s.advise(1e-3)
        assert c.d[GRAPHS] == 1 and c.d[EAGER] == 0                            # the entry kept its graph
    print("ok read_in_the_middle")


def scenario_periodic_patterns():
    """Bulirsch-Stoer-like: 2 sub-steps of one size, then 4 of another, per outer step; and an A B B B pattern."""
    y, m = system()
    for pattern in ([1e-3] * 2 + [5e-4] * 4, [1e-3, 5e-4, 5e-4, 5e-4], [1e-3, 5e-4]):
        with Engine() as e:
            assert e.init(y, m)
            e.set_option("step_graph", 1)
            s = Rk4(e)
            for _ in range(6):
                for dt in pattern:
                    s.advise(dt)
            bail0 = e.step_graph_stats()["bailouts"]
            with Delta() as c:
                for _ in range(5):
                    for dt in pattern:
                        s.advise(dt)
            st = e.step_graph_stats()
            assert st["distinct_steps"] == 2
            assert c.d[CAPTURES] == 0 and c.d[INSTANTIATED] == 0               # learnt: nothing is captured any more
            steps = 5 * len(pattern)
            wrong = st["bailouts"] - bail0                                     # a wrong guess seen only at the 2nd call costs that step
            assert c.d[GRAPHS] == steps - wrong and c.d[EAGER] == 8 * wrong
            assert wrong <= 5 * 2
    print("ok periodic_patterns")


def scenario_never_repeating_and_table_limit():
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        s = Rk4(e)
        with Delta() as c:
            for i in range(100):
                s.advise(1e-3 * (1 + 1e-3 * i))
        st = e.step_graph_stats()
        assert c.d[CAPTURES] == 0 and c.d[GRAPHS] == 0 and c.d[EAGER] == 800   # never predicted: plain eager execution
        assert st["state"] == "record" and 1 <= st["distinct_steps"] <= 64     # the table starts over when it is full
    print("ok never_repeating_and_table_limit")


def scenario_buffers_change_empties_the_table():
    y, m = system()
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        s = Rk4(e)
        for _ in range(5):
            s.advise(1e-3)
        assert e.step_graph_stats()["state"] == "replay"
        extra = e.create_buffer(64)                                             # a new handle may reuse an old address
        st = e.step_graph_stats()
        assert st["state"] == "record" and st["distinct_steps"] == 0
        with Delta() as c:
            s.advise(1e-3)
        assert c.d[EAGER] == 8
        e.free_buffer(extra)
        for _ in range(4):
            s.advise(1e-3)
        assert e.step_graph_stats()["state"] == "replay"
        with Engine(devices="0,0") as lanes:                                    # several shards: accepted and ignored
            assert lanes.init(y, m)
            assert lanes.set_option("step_graph", 1) == 0 and lanes.step_graph_stats()["state"] == "off"
    print("ok buffers_change_empties_the_table")


def scenario_tree_build_rate():
    """Barnes-Hut with tree_build_rate = 4: every 4th step rebuilds the tree (3 radix sorts + partitions + local builds),
    the others only refresh it. Two table entries; the rebuild step is captured on the fly the first time it arrives where
    a refresh step was predicted, and from then on every step of either kind is one graph launch."""
    n = 4096
    y, m = system(n)
    with Engine(kind="bh", tree_build_rate=4) as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        f = e.create_buffer(y.nbytes)
        kernels = []
        for i in range(20):
            with Delta() as c:
                e.fcompute(0.0, e.get_y(), f)
                e.fmadd_inplace(e.get_y(), f, 1e-3)
                e.advise_time(1e-3)
            kernels.append(int(c.d[EAGER] + c.d[REPLAYED]))
            if i >= 8:
                assert c.d[GRAPHS] == 1 and c.d[EAGER] == 0 and c.d[CAPTURES] == 0
        st = e.step_graph_stats()
        assert st["distinct_steps"] == 2 and st["bailouts"] == 0
        rebuild, refresh = kernels[0], kernels[1]
        assert rebuild > refresh and kernels == [rebuild if i % 4 == 0 else refresh for i in range(20)]
    print("ok tree_build_rate")


def scenario_direct_path_selection():
    """Which kernel family nb200_fcompute_direct picks (nb200_last_direct_path: -1 single launch, 0 ordered pairs, else the
    symmetric tile edge) and how many kernels that is, by N, precision, shard count and tunables."""
    cases = [("f64", "0", 16, (), -1, 1), ("f64", "0", 2048, (), -1, 1), ("f64", "0", 4096, (), -1, 1),
             ("f64", "0", 4160, (), 0, 3),                                       # pack + pairs + reduce
             ("f64", "0", 8192, (), 256, 4), ("f64", "0", 32768, (), 256, 4),  # pack + tiles + reduce + finish
             ("f64", "0", 65536, (), 512, 4), ("f64", "0", 131072, (), 1024, 4), ("f64", "0", 262144, (), 2048, 4),
             ("f32", "0", 8192, (), 0, 3), ("f32", "0", 16384, (), 256, 4), ("f32", "0", 65536, (), 512, 4),
             ("f64", "0,0", 2048, (), 0, 6),                                    # several shards: the tiled path
             ("f64", "0,0", 65536, (), 512, 10),                                # + peer sum of the partials, per lane
             ("f64", "0", 2048, (("direct_small", 0),), 0, 3), ("f64", "0", 2048, (("direct_symmetric", 0),), 0, 3),
             ("f64", "0", 65536, (("direct_symmetric", 0),), 0, 3), ("f64", "0", 2048, (("direct_symmetric", 1),), 256, 4),
             ("f64", "0", 6000, (("direct_small", 1),), -1, 1),
             ("f64", "0", 65536, (("direct_sym_tile", 1536),), 0, 3)]           # not a power of two: ordered pairs
    for precision, devices, n, opts, want_path, want_launches in cases:
        y, m = system(n)
        with Engine(precision=precision, devices=devices) as e:
            for k, v in opts:
                e.set_option(k, v)
            assert e.init(y, m)
            f = e.create_buffer(e.get_y().size())
            e.fcompute(0.0, e.get_y(), f)                                       # first call: scratch allocation
            before = e.launch_count()
            with Delta() as c:
                e.fcompute(0.0, e.get_y(), f)
            got = (e.last_direct_path(), e.launch_count() - before)
            assert got == (want_path, want_launches) and c.d[EAGER] == want_launches, (precision, devices, n, opts, got)
    print("ok direct_path_selection")


def scenario_bodies_and_statistics_are_host_visible():
    y, m = system(256)
    n = m.size
    pos, vel = np.ascontiguousarray(y.reshape(6, n)[:3].T), np.ascontiguousarray(y.reshape(6, n)[3:].T)
    for devices, lanes in (("0", 1), ("0,0,0,0", 4)):
        with Engine(devices=devices) as e:
            with Delta() as c:
                assert e.init_bodies(pos, vel, m)
            assert c.d[EAGER] == lanes and c.d[COPIES] >= 2 * lanes             # two uploads + one transpose kernel per lane
            with Delta() as c:
                assert e.get_bodies() is not None
            assert c.d[EAGER] == lanes and c.d[COPIES] == 2 * lanes             # one transpose kernel + two downloads per lane
            assert e.host_register(pos) == 0 and e.host_unregister(pos) == 0
            small = e.create_buffer(64)
            assert e.lib.nb200_read_bodies(e.ctx, small.handle, pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p)) == -1
            assert e.lib.nb200_read_bodies(e.ctx, e.get_y().handle, None, None) == -1
    with Engine() as e:
        assert e.init(y, m)
        e.set_option("step_graph", 1)
        dy = e.create_buffer(y.nbytes)
        for _ in range(5):
            e.fcompute(0.0, e.get_y(), dy)
            e.fmadd_inplace(e.get_y(), dy, 1e-3)
            e.advise_time(1e-3)
        assert e.step_graph_stats()["state"] == "replay"
        e.fcompute(0.0, e.get_y(), dy)
        with Delta() as c:
            e.statistics(with_energy=False)                                     # host-visible: the accepted fcompute runs first
        assert c.d[EAGER] >= 2 and e.step_graph_stats()["bailouts"] == 1
    print("ok bodies_and_statistics_are_host_visible")


if __name__ == "__main__":
    scenario_direct_path_selection()
    scenario_bodies_and_statistics_are_host_visible()
    scenario_tree_build_rate()
    scenario_buffers_and_shards()
    scenario_fixed_step_replay()
    scenario_deferral_is_real()
    scenario_error_norm_border()
    scenario_read_in_the_middle()
    scenario_periodic_patterns()
    scenario_never_repeating_and_table_limit()
    scenario_buffers_change_empties_the_table()
    print("all host-logic scenarios passed")
