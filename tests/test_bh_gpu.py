"""Barnes-Hut on the GPU vs the reference's nbody_space_heap / simple_bh (tests/golden/*.npz, generated from
oracle/_ref) and the CPU oracle: identical tree (leaf order exact, node data to rounding) and per-body
acceleration <= 1e-12 relative, mirroring the cuda_bh_tex rows of test_nbody_engine.cpp:1227-1336."""
import numpy as np
import pytest

from conftest import load_golden_npz, golden_path, rel_err_per_body
from util import universe

pytestmark = pytest.mark.gpu
TOL = 1e-12


def run_bh(y, m, ratio, precision="f64", devices="0", layout="heap_stackless", rate=0, walk_mode=0, steps=1, stats=False,
           options=()):
    """Returns list of f per step (y *= 0.99 between steps, like test_fcompute), the exported tree and stats."""
    from nbody_b200 import Engine
    out = []
    with Engine(precision=precision, devices=devices, kind="bh", distance_to_node_radius_ratio=ratio,
                tree_layout=layout, tree_build_rate=rate) as e:
        e.set_option("walk_mode", walk_mode)
        for k, v in options:
            e.set_option(k, v)
        assert e.init(y, m)
        if stats:
            e.bh_walk_stats(True)
        ybuf = e.create_buffer(e.get_y().size())
        e.copy_buffer(ybuf, e.get_y())
        f = e.create_buffer(e.get_y().size())
        for step in range(steps):
            e.fill_buffer(f, -1e10)
            e.set_step(step)
            e.fcompute(0.0, ybuf, f)
            out.append(e.read_buffer(f))
            e.fmadd_inplace(ybuf, ybuf, -0.01)
        tree = e.bh_export_tree()
        st = e.bh_walk_stats(True) if stats else None
    return out, tree, st


@pytest.mark.parametrize("tag,n,ratio,key", [("g1_n128", 128, 3.1623, "3p1623"), ("g1_n128", 128, 10.0, "10"),
                                             ("g1_n256", 256, 3.1623, "3p1623"), ("g1_n256", 256, 10.0, "10"),
                                             ("g1_n2048", 2048, 10.0, "10")])
@pytest.mark.parametrize("layout", ["heap", "heap_stackless"])
def test_bh_vs_reference_simple_bh(tag, n, ratio, key, layout):
    g = load_golden_npz(tag)
    (f,), (xyzr, mass, body), _ = run_bh(g["y"], g["mass"], ratio, layout=layout)
    assert np.array_equal(body[n:], g["tree_body_" + key][n:])                 # same leaf order as nbody_space_heap::build
    assert np.allclose(mass[1:], g["tree_mass_" + key][1:], rtol=1e-15, atol=0)
    assert np.allclose(xyzr[1:, :3], g["tree_xyzr_" + key][1:, :3], rtol=1e-13, atol=1e-13)
    assert np.allclose(xyzr[1:, 3], g["tree_xyzr_" + key][1:, 3], rtol=1e-11, atol=0)
    assert np.array_equal(f[:3 * n], g["y"][3 * n:])
    assert rel_err_per_body(f, g["f_bh_" + key], n) <= TOL


def test_bh_huge_ratio_equals_direct():
    """ratio 1e8 opens every internal node: BH == direct sum (reference gate 1e-11 absolute)."""
    g = load_golden_npz("g1_n256")
    (f,), _, st = run_bh(g["y"], g["mass"], 1e8, stats=True)
    assert np.abs(f - g["f_simple"]).max() <= 1e-11
    assert rel_err_per_body(f, g["f_bh_1e8"], 256) <= TOL
    assert st[1] == 256 * 255


def test_bh_walk_modes_bit_identical():
    """The walks that keep one traversal state per target (thread per target, one / two / four targets per lane) add a
    target's accepted nodes in the order of nbody_space_heap_stackless::traverse: bit-identical to each other."""
    g = load_golden_npz("g1_n2048")
    (a,), _, sa = run_bh(g["y"], g["mass"], 10.0, walk_mode=2, stats=True)
    for mode in (1, 32):
        (b,), _, sb = run_bh(g["y"], g["mass"], 10.0, walk_mode=mode, stats=True)
        assert np.array_equal(a, b)
        assert sa == sb and sa[0] > sa[1] > 0


GROUP_TOL = {"f64": 1e-13, "f32": 3e-4}


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("n,ratio,devices", [(2, 10.0, "0"), (64, 10.0, "0"), (2048, 10.0, "0"), (2048, 1.0, "0,0"),
                                             (65536, 10.0, "0"), (65536, 2.0, "0,0,0,0")])
def test_bh_several_targets_per_lane_bit_identical(precision, n, ratio, devices):
    """walk_mode 2 / 4: one warp walks the union of 64 / 128 consecutive leaves (2 / 4 targets per lane); every target
    still accepts exactly the nodes of its own stackless traversal, so forces and visit counts equal walk_mode 32 (one
    target per lane). walk_mode 0 / 8, the default, is the grouped walk (nb200_bh_group.cuh): the same accepted nodes per
    target (equal visit and interaction counts), added in another order -- equal to rounding, not bit for bit."""
    y, m = universe(n, precision) if n >= 128 else (None, None)   # make_universe rounds up to 2 x 64 bodies
    if y is None:
        g = load_golden_npz("g1_n128", precision)
        idx = np.arange(n)
        y = np.concatenate([g["y"][r * 128 + idx] for r in range(6)])
        m = g["mass"][idx]
    (a,), _, sa = run_bh(y, m, ratio, precision=precision, devices=devices, walk_mode=32, stats=True)
    for mode in (2, 4):
        (b,), _, sb = run_bh(y, m, ratio, precision=precision, devices=devices, walk_mode=mode, stats=True)
        assert np.array_equal(a, b), "walk_mode %d" % mode
        assert sa == sb
    for mode in (0, 8):
        (b,), _, sb = run_bh(y, m, ratio, precision=precision, devices=devices, walk_mode=mode, stats=True)
        assert sa == sb, "walk_mode %d: visits / interactions %r != %r" % (mode, sb, sa)
        assert np.array_equal(a[:3 * n], b[:3 * n])
        assert rel_err_per_body(b, a, n) <= GROUP_TOL[precision], "walk_mode %d" % mode


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_bh_grouped_walk_deterministic_and_shard_independent(precision):
    """The grouped walk's summation order is a pure function of the inputs: two runs agree bit for bit, and so do runs
    with 1, 2 and 4 shards (a group is the same 32 consecutive leaves however the chunks are dealt)."""
    y, m = universe(65536, precision)
    (a,), _, _ = run_bh(y, m, 10.0, precision=precision)
    (b,), _, _ = run_bh(y, m, 10.0, precision=precision)
    assert np.array_equal(a, b)
    for devices in ("0,0", "0,0,0,0"):
        (c,), _, _ = run_bh(y, m, 10.0, precision=precision, devices=devices)
        assert np.array_equal(a, c), devices


def test_bh_grouped_walk_knife_edge_counts(oracle64):
    """Certified FP32 decisions: bodies on a lattice with equal masses put many (target, node) pairs exactly on
    d2 == radius_sqr; every such test must fall back to the FP64 expression and decide as simple_bh does."""
    k = 16
    g = np.arange(k, dtype=np.float64)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)      # 4096 lattice points, spacing 1
    n = pos.shape[0]
    y = np.concatenate([pos[:, 0], pos[:, 1], pos[:, 2], np.zeros(3 * n)])
    m = np.ones(n)
    for ratio in (1.0, 2.0, 10.0):
        t = oracle64.heap_build(y, m, ratio)
        want, visits, inter = oracle64.fcompute_bh(y, m, t)
        (f,), _, st = run_bh(y, m, ratio, stats=True)
        assert st == (visits, inter), ratio
        assert rel_err_per_body(f, want, n, floor=1e-9) <= 1e-12, ratio


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_bh_longest_walk_first_order_changes_nothing(devices):
    """walk_lpt: from the second walk on the CTAs are launched most-expensive-first (costs of the previous walk, sorted
    on the device). Which CTA walks which leaves does not change any target's traversal: bit-identical forces."""
    y, m = universe(65536)
    a, _, sa = run_bh(y, m, 10.0, devices=devices, steps=3, stats=True, options=(("walk_lpt", 0),))
    b, _, sb = run_bh(y, m, 10.0, devices=devices, steps=3, stats=True, options=(("walk_lpt", 1),))
    assert sa == sb
    for fa, fb in zip(a, b):
        assert np.array_equal(fa, fb)


def test_bh_counts_match_oracle(oracle64):
    g = load_golden_npz("g1_n2048")
    t = oracle64.heap_build(g["y"], g["mass"], 10.0)
    _, visits, inter = oracle64.fcompute_bh(g["y"], g["mass"], t)
    _, _, st = run_bh(g["y"], g["mass"], 10.0, stats=True)
    assert st == (visits, inter)


@pytest.mark.parametrize("rate", [0, 2])
def test_bh_tree_build_rate(oracle64, rate):
    """tree_build_rate {0, 2} x 2 steps (test_nbody_engine.cpp:1283-1336): with rate 2 the second step only
    refreshes the geometry of the step-0 topology -- compared with the oracle doing the same."""
    g = load_golden_npz("g1_n256")
    fs, _, _ = run_bh(g["y"], g["mass"], 3.1623, rate=rate, steps=3)
    y = g["y"].copy()
    tree = None
    for step in range(3):
        if rate == 0 or tree is None or step % rate == 0:
            tree = oracle64.heap_build(y, g["mass"], 3.1623)
        else:
            tree = oracle64.heap_rebuild(tree, y)
        want, _, _ = oracle64.fcompute_bh(y, g["mass"], tree)
        assert rel_err_per_body(fs[step], want, 256) <= TOL, step
        y = y + y * -0.01


def test_bh_n16_golden_state(oracle64):
    from oracle.oracle import load_table
    y, m = load_table(golden_path("initial_state.txt"))
    (f,), (xyzr, mass, body), _ = run_bh(y, m, 10.0)
    t = oracle64.heap_build(y, m, 10.0)
    assert np.array_equal(body[16:], t["body_n"][16:].astype(np.int32))
    want, _, _ = oracle64.fcompute_bh(y, m, t)
    assert rel_err_per_body(f, want, 16) <= TOL


@pytest.mark.parametrize("n", [2, 4, 1024, 4096])
def test_bh_small_and_boundary_sizes(oracle64, n):
    """N = 1024 is the last single-CTA build, N = 4096 the first with two partition levels."""
    rng = np.random.RandomState(n)
    y = rng.uniform(-10, 10, 6 * n)
    m = rng.uniform(0.1, 2.0, n)
    (f,), (_, _, body), _ = run_bh(y, m, 2.0)
    t = oracle64.heap_build(y, m, 2.0)
    assert np.array_equal(body[n:], t["body_n"][n:].astype(np.int32))
    want, _, _ = oracle64.fcompute_bh(y, m, t)
    assert rel_err_per_body(f, want, n) <= TOL


def test_bh_rejects_non_power_of_two():
    from nbody_b200 import Engine
    rng = np.random.RandomState(0)
    with Engine(kind="bh") as e:
        assert e.init(rng.rand(6 * 12), np.ones(12))
        f = e.create_buffer(e.get_y().size())
        e.fill_buffer(f, 7.0)
        e.fcompute(0.0, e.get_y(), f)
        assert np.all(e.read_buffer(f) == 7.0)
        assert "power of two" in e.last_error()


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0"])
def test_bh_sharded_lanes_equal_single(devices):
    g = load_golden_npz("g1_n2048")
    (one,), _, _ = run_bh(g["y"], g["mass"], 10.0)
    (many,), _, _ = run_bh(g["y"], g["mass"], 10.0, devices=devices)
    assert np.array_equal(one, many)


def test_bh_fp32():
    g = load_golden_npz("g1_n2048", "f32")
    (f,), (_, _, body), _ = run_bh(g["y"], g["mass"], 10.0, precision="f32")
    assert np.array_equal(body[2048:], g["tree_body_10"][2048:])
    assert rel_err_per_body(f, g["f_bh_10"], 2048) <= 1e-4


def test_bh_n65536_full_vs_oracle(oracle64):
    """Phase-A build (6 partition levels) + walk at N = 65,536, ratio 3.1623, every body checked."""
    n = 65536
    y, m = universe(n)
    (f,), (xyzr, mass, body), st = run_bh(y, m, 3.1623, stats=True)
    t = oracle64.heap_build(y, m, 3.1623)
    assert np.array_equal(body[n:], t["body_n"][n:].astype(np.int32))
    assert np.allclose(xyzr[1:, :3], t["xyzr"][1:, :3], rtol=1e-12, atol=1e-12)
    want, visits, inter = oracle64.fcompute_bh(y, m, t)
    assert st == (visits, inter)
    assert rel_err_per_body(f, want, n) <= TOL


def test_bh_n1m_tree_equals_oracle_and_walk_ratio1(oracle64):
    """N = 1,048,576: leaf order identical to the CPU nth_element build; walk at ratio 1 vs the oracle."""
    n = 1 << 20
    y, m = universe(n)
    (f,), (_, _, body), st = run_bh(y, m, 1.0, stats=True)
    t = oracle64.heap_build(y, m, 1.0)
    assert np.array_equal(body[n:], t["body_n"][n:].astype(np.int32))
    want, visits, inter = oracle64.fcompute_bh(y, m, t)
    assert st == (visits, inter)
    assert rel_err_per_body(f, want, n) <= TOL


def test_bh_c4_n4m_ratio10_sampled(oracle64):
    """BASELINE config C4 (N = 4,194,304, ratio 10): leaf order identical to the CPU build; accelerations of 384
    sampled targets (incl. both central bodies' leaves) vs the oracle's stackless walk of the same tree."""
    n = 1 << 22
    y, m = universe(n)
    (f,), (xyzr, mass, body), _ = run_bh(y, m, 10.0)
    t = oracle64.heap_build(y, m, 10.0)
    assert np.array_equal(body[n:], t["body_n"][n:].astype(np.int32))
    assert np.allclose(xyzr[1:, :3], t["xyzr"][1:, :3], rtol=1e-12, atol=1e-12)
    leaf_of = np.empty(n, dtype=np.int64)
    leaf_of[t["body_n"][n:]] = np.arange(n)
    bodies = np.unique(np.concatenate([[0, n // 2, n - 1], np.random.RandomState(4).randint(0, n, 381)]))
    want = oracle64.bh_subset(t, leaf_of[bodies])
    got = f.reshape(6, n)[3:, bodies]
    assert rel_err_per_body(got, want, bodies.size) <= TOL
    assert np.array_equal(f[:3 * n], y[3 * n:])


def test_bh_walk_profile_counters_are_consistent():
    """nb200_bh_walk_profile: rounds of at most 32 work items, at most 64 list entries per round, every entry naming at
    least one target; the entry histogram covers all entries but each group's last round."""
    from nbody_b200 import Engine
    y, m = universe(65536)
    with Engine(kind="bh", distance_to_node_radius_ratio=10.0) as e:
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.bh_walk_stats(True)
        e.fcompute(0.0, e.get_y(), f)
        e.synchronize()
        p = e.bh_walk_profile()
        visits, inter = e.bh_walk_stats(False)
    groups = 65536 // 32
    assert 0 < p["items"] <= 32 * p["rounds"] and p["entries"] <= 64 * p["rounds"] + groups
    assert visits == 2 * 0 + visits and visits >= 2 * p["items"]            # every item is two node visits by >= 1 target
    assert p["entries"] <= inter <= 32 * p["entries"]
    hist = sum(p[k] for k in ("entries_32", "entries_24_31", "entries_16_23", "entries_8_15", "entries_1_7"))
    assert p["entries"] - 64 * groups <= hist <= p["entries"]
    assert p["busiest_target_entries"] <= p["entries"] and p["busiest_target_entries_whole_walk"] * 32 >= inter * 0.9
    assert 0 < p["max_stack"] <= 512 and p["unsure_lane_items"] < p["items"] // 100


def test_bh_beyond_the_headline_size_and_a_tiny_stack_budget():
    """N = 8,388,608 (one level deeper than C4), ratio 4: the grouped walk against the two-targets-per-lane walk, whose
    order of summation is the reference's -- equal node-visit and interaction counts, forces to rounding -- and the
    item stack stays well inside its 512 slots (above 448 the walk would fall back to one item at a time)."""
    from nbody_b200 import Engine
    n = 1 << 23
    y, m = universe(n)
    out = {}
    for mode in (2, 0):
        with Engine(kind="bh", distance_to_node_radius_ratio=4.0) as e:
            e.set_option("walk_mode", mode)
            assert e.init(y, m)
            f = e.create_buffer(e.get_y().size())
            e.bh_walk_stats(True)
            e.fcompute(0.0, e.get_y(), f)
            e.synchronize()
            prof = e.bh_walk_profile() if mode == 0 else None
            out[mode] = (e.read_buffer(f), e.bh_walk_stats(False), prof)
    assert out[0][1] == out[2][1] and out[0][1][0] > out[0][1][1] > 0
    # two orders of summation over ~1e4 terms per target: equal to rounding. Among 8M bodies a few have attractions that
    # nearly cancel, which amplifies the relative figure; measured against the typical size of an acceleration it is 1e-13
    assert rel_err_per_body(out[0][0], out[2][0], n) <= 1e-12
    assert rel_err_per_body(out[0][0], out[2][0], n, floor=1e-2) <= 1e-13
    assert 0 < out[0][2]["max_stack"] < 448
