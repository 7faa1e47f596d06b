"""Parity of the CUDA direct all-pairs fcompute (through the C ABI) with the
reference's engines / the CPU oracle. Gate: per-body relative error of the
acceleration <= 1e-12 in FP64 (BASELINE.json north_star), ~1e-5 in FP32."""
import numpy as np
import pytest

from conftest import load_golden_npz, golden_path, rel_err_per_body
from util import universe

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-4     # vs the reference's own FP32 result, whose sequential FP32 sum carries ~1e-5 of noise itself


def run_direct(y, m, precision="f64", devices="0", options=()):
    from nbody_b200 import Engine
    with Engine(precision=precision, devices=devices) as e:
        for k, v in options:
            e.set_option(k, v)
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.fill_buffer(f, -1e10)
        e.fcompute(0.0, e.get_y(), f)
        out = e.read_buffer(f)
        assert e.get_compute_count() == 1 and e.launch_count() >= 2
        e.free_buffer(f)
    return out


# Up to N = 4,096 on one shard the automatic choice is the single-launch kernel (direct_small); direct_small=0 selects
# the tiled ordered-pair kernel (pack + direct_pairs + direct_reduce) that larger and sharded systems use.
SMALL_PATHS = pytest.mark.parametrize("path", [(), (("direct_small", 0),)], ids=["single-launch", "ordered-pairs"])


@SMALL_PATHS
@pytest.mark.parametrize("tag,n", [("g1_n128", 128), ("g1_n256", 256), ("g1_n2048", 2048)])
def test_direct_vs_reference_engines_fp64(tag, n, path):
    g = load_golden_npz(tag)
    f = run_direct(g["y"], g["mass"], options=path)
    assert np.array_equal(f[:3 * n], g["y"][3 * n:])          # dr/dt = v, bit-exact
    assert rel_err_per_body(f, g["f_openmp"], n) <= TOL64
    assert rel_err_per_body(f, g["f_block"], n) <= TOL64
    # the reference's own cross-engine gate: |df| <= 1e-13 absolute vs nbody_engine_simple
    # (test_nbody_engine.cpp:472-558, default m_eps) holds for the accelerations of this fixture
    assert np.abs(f - g["f_simple"]).max() <= 1e-13 * max(1.0, np.abs(g["f_simple"]).max())


@SMALL_PATHS
@pytest.mark.parametrize("tag,n", [("g1_n128", 128), ("g1_n2048", 2048)])
def test_direct_vs_reference_engines_fp32(oracle64, tag, n, path):
    g = load_golden_npz(tag, "f32")
    f = run_direct(g["y"], g["mass"], precision="f32", options=path)
    assert np.array_equal(f[:3 * n], g["y"][3 * n:])
    assert rel_err_per_body(f, g["f_openmp"], n) <= TOL32
    # against FP64 arithmetic on the same FP32 inputs the GPU result is at least as accurate as the reference's
    truth = oracle64.fcompute_openmp(g["y"].astype(np.float64), g["mass"].astype(np.float64))
    err_gpu = rel_err_per_body(f, truth, n)
    err_ref = rel_err_per_body(g["f_openmp"], truth, n)
    assert err_gpu <= max(2 * err_ref, 2e-6), (err_gpu, err_ref)


@SMALL_PATHS
def test_direct_n16_golden_state(oracle64, path):
    """N = 16 (the solver golden state): not a multiple of any tile size."""
    from oracle.oracle import load_table
    y, m = load_table(golden_path("initial_state.txt"))
    f = run_direct(y, m, options=path)
    assert rel_err_per_body(f, oracle64.fcompute_openmp(y, m), 16) <= TOL64


@SMALL_PATHS
@pytest.mark.parametrize("n", [1, 2, 3, 15, 17, 129, 1000, 4096])
def test_direct_ragged_sizes(oracle64, n, path):
    rng = np.random.RandomState(n)
    y = rng.uniform(-10, 10, 6 * n)
    m = rng.uniform(0.1, 2.0, n)
    f = run_direct(y, m, options=path)
    assert np.array_equal(f[:3 * n], y[3 * n:])
    ref = oracle64.fcompute_openmp(y, m)
    if n == 1:
        assert np.all(f[3:] == 0)
    else:
        assert rel_err_per_body(f, ref, n) <= TOL64


@SMALL_PATHS
def test_direct_coincident_bodies_use_min_distance(oracle64, path):
    """r^2 < 1e-8 is clamped to 1e-8 (nbody_data.cpp:39-42); coincident bodies contribute exactly 0."""
    y = np.zeros(6 * 4)
    y[0:4] = [0.0, 0.0, 5e-5, 1.0]          # bodies 0 and 1 coincide; body 2 is 5e-5 away (r^2 = 2.5e-9 < 1e-8)
    m = np.array([1.0, 2.0, 3.0, 4.0])
    f = run_direct(y, m, options=path)
    ref = oracle64.fcompute_openmp(y, m)
    assert np.all(np.isfinite(f))
    assert np.allclose(f, ref, rtol=1e-13, atol=0)


@pytest.mark.parametrize("ipt,segments", [(1, 1), (2, 3), (4, 1), (4, 16)])
def test_direct_all_kernel_shapes(ipt, segments):
    """Every (targets per thread) x (source segments) shape gives the same physics; shape (4, 1) writes f directly."""
    g = load_golden_npz("g1_n2048")
    f = run_direct(g["y"], g["mass"], options=(("direct_targets_per_thread", ipt), ("direct_segments", segments)))
    assert rel_err_per_body(f, g["f_openmp"], 2048) <= TOL64
    assert np.array_equal(f[:3 * 2048], g["y"][3 * 2048:])


@SMALL_PATHS
def test_direct_is_deterministic(path):
    g = load_golden_npz("g1_n2048")
    a = run_direct(g["y"], g["mass"], options=path)
    b = run_direct(g["y"], g["mass"], options=path)
    assert np.array_equal(a, b)


def test_direct_single_launch_kernel_is_the_small_n_default(oracle64):
    """C1 (N = 2,048): fcompute is ONE kernel launch (pairs + slice reduction + velocity rows); above N = 4,096, with
    several shards, or with a tiled-path tunable set, the tiled kernels run. Forced (direct_small=1) it handles any N."""
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n2048")
    for devices, opts, want_path, want_launches in (("0", (), -1, 1), ("0", (("direct_small", 0),), 0, 3), ("0,0", (), 0, 6),
                                                    ("0", (("direct_segments", 4),), 0, 3)):
        with Engine(devices=devices) as e:
            for k, v in opts:
                e.set_option(k, v)
            assert e.init(g["y"], g["mass"])
            f = e.create_buffer(e.get_y().size())
            before = e.launch_count()
            e.fcompute(0.0, e.get_y(), f)
            assert e.last_direct_path() == want_path
            assert e.launch_count() - before == want_launches
            assert rel_err_per_body(e.read_buffer(f), g["f_openmp"], 2048) <= TOL64
    n = 6000
    rng = np.random.RandomState(6)
    y, m = rng.uniform(-50, 50, 6 * n), rng.uniform(0.1, 2.0, n)
    f = run_direct(y, m, options=(("direct_small", 1),))
    assert rel_err_per_body(f, oracle64.fcompute_openmp(y, m), n) <= TOL64


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0"])
def test_direct_sharded_lanes_equal_single(devices):
    """The reference's duplicate-device trick (test_nbody_engine.cpp:1259-1267): body-sharded lanes on one GPU.
    Its gate for cuda vs multi-device cuda is 1e-15; shards sum the same sources in the same order."""
    g = load_golden_npz("g1_n2048")
    one = run_direct(g["y"], g["mass"], options=(("direct_segments", 4), ("direct_targets_per_thread", 1)))
    many = run_direct(g["y"], g["mass"], devices=devices, options=(("direct_segments", 4), ("direct_targets_per_thread", 1)))
    assert np.array_equal(one, many)


def test_direct_c2_n65536_sampled(oracle64):
    """BASELINE config C2 (N = 65,536): GPU result vs the oracle on a 256-body sample incl. both central bodies."""
    n = 65536
    y, m = universe(n)
    f = run_direct(y, m).reshape(6, n)
    t = np.unique(np.concatenate([[0, n // 2, n - 1], np.random.RandomState(3).randint(0, n, 253)]))
    ref = oracle64.accel_subset(y, m, t)
    assert rel_err_per_body(f[3:, t], ref, t.size) <= TOL64
    ld = oracle64.accel_subset(y, m, t, long_double=True)
    assert rel_err_per_body(f[3:, t], ld, t.size) <= TOL64
    assert np.array_equal(f[:3].reshape(-1), y[3 * n:])


def test_direct_c3_n1m_sampled_and_momentum(oracle64):
    """BASELINE config C3 (N = 1,048,576): 64-body sample vs the oracle, plus a size-independent property:
    total force sum_i m_i a_i = 0 (Newton's third law) to rounding."""
    n = 1 << 20
    y, m = universe(n)
    f = run_direct(y, m).reshape(6, n)
    t = np.unique(np.concatenate([[0, n // 2], np.random.RandomState(5).randint(0, n, 62)]))
    ref = oracle64.accel_subset(y, m, t)
    assert rel_err_per_body(f[3:, t], ref, t.size) <= TOL64
    force = (f[3:] * m[None, :]).sum(axis=1)
    scale = np.abs(f[3:] * m[None, :]).sum(axis=1)
    assert np.all(np.abs(force) <= 1e-10 * scale)


def test_direct_n4m_uses_the_symmetric_tiles_and_stays_exact(oracle64):
    """Four times BASELINE's C3: N = 4,194,304 (1.76e13 pairs, 51 GB of tile partials on the 180 GB part) still takes
    the symmetric tiles; 24 sampled bodies vs the oracle and Newton's third law over all of them."""
    from nbody_b200 import Engine
    n = 1 << 22
    y, m = universe(n)
    with Engine() as e:
        assert e.init(y, m)
        fb = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), fb)
        f = e.read_buffer(fb).reshape(6, n)
        assert e.last_direct_path() == 8192
    t = np.unique(np.concatenate([[0, n // 2], np.random.RandomState(6).randint(0, n, 22)]))
    ref = oracle64.accel_subset(y, m, t)
    assert rel_err_per_body(f[3:, t], ref, t.size) <= TOL64
    force = (f[3:] * m[None, :]).sum(axis=1)
    scale = np.abs(f[3:] * m[None, :]).sum(axis=1)
    assert np.all(np.abs(force) <= 1e-10 * scale)


def test_multi_process_nccl_two_ranks():
    """One process per GPU over NCCL (skipped on a single-GPU box; the gloo tests cover the host logic there)."""
    import os
    import subprocess
    import sys
    from nbody_b200 import device_count
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "mp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "mp_check direct ok" in out.stdout and "mp_check bh ok" in out.stdout and "mp_check direct-symmetric ok" in out.stdout


# ---- symmetric (Newton's third law) tile path: automatic for N >= 8,192, forced here on small systems ------------
@pytest.mark.parametrize("n,tile,shape", [(4096, 1024, 0), (8192, 2048, 0), (5000, 1024, 0), (4096, 1024, 1), (2048, 256, 0),
                                          (3000, 512, 1), (2048, 256, 1), (1000, 256, 1), (4096, 512, 1)])
def test_direct_symmetric_tiles_vs_oracle(oracle64, n, tile, shape):
    """Every unordered pair evaluated once and applied to both bodies; ragged N pads the last tile with zero-mass
    bodies; diagonal tiles are one-sided. Same 1e-12 gate as the plain kernel."""
    rng = np.random.RandomState(n + tile)
    y = rng.uniform(-50, 50, 6 * n)
    m = rng.uniform(0.1, 2.0, n)
    y[0:2] = y[2]                      # bodies 0, 1, 2 share an x coordinate ...
    y[n:n + 2] = y[n + 2]
    y[2 * n] = y[2 * n + 2]            # ... and bodies 0 and 2 coincide entirely (clamped pair)
    y[2 * n + 1] = y[2 * n + 2] + 5e-5
    opts = (("direct_symmetric", 1), ("direct_sym_tile", tile), ("direct_sym_shape", shape))
    f = run_direct(y, m, options=opts)
    ref = oracle64.fcompute_openmp(y, m)
    assert np.array_equal(f[:3 * n], y[3 * n:])
    assert np.all(np.isfinite(f))
    assert rel_err_per_body(f, ref, n) <= TOL64
    plain = run_direct(y, m, options=(("direct_symmetric", 0),))
    assert rel_err_per_body(f, plain, n) <= 1e-13
    again = run_direct(y, m, options=opts)
    assert np.array_equal(f, again)                     # fixed summation order: bit-reproducible


@pytest.mark.parametrize("devices,n,tile", [("0,0", 4096, 1024), ("0,0,0", 6144, 512), ("0,0,0,0", 8192, 1024)])
def test_direct_symmetric_tiles_across_lanes(oracle64, devices, n, tile):
    """Lanes of one process deal the tiles round-robin; each lane sums its body block of every lane's partial vector
    straight from the peers' memory, lanes in ascending order. Two fcomputes back to back exercise the
    'peers finished reading my partials' events."""
    rng = np.random.RandomState(n)
    y = rng.uniform(-50, 50, 6 * n)
    m = rng.uniform(0.1, 2.0, n)
    opts = (("direct_symmetric", 1), ("direct_sym_tile", tile))
    from nbody_b200 import Engine
    with Engine(devices=devices) as e:
        for k, v in opts:
            e.set_option(k, v)
        assert e.init(y, m)
        f1 = e.create_buffer(e.get_y().size())
        f2 = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), f1)
        e.fcompute(0.0, f1, f2)            # different input, same scratch, no host sync in between
        e.fcompute(0.0, e.get_y(), f2)
        assert e.last_direct_path() == tile
        a, b = e.read_buffer(f1), e.read_buffer(f2)
    assert np.array_equal(a, b)
    ref = oracle64.fcompute_openmp(y, m)
    assert rel_err_per_body(a, ref, n) <= TOL64
    assert rel_err_per_body(a, run_direct(y, m, options=opts), n) <= 1e-13


def test_direct_symmetric_golden_universe():
    g = load_golden_npz("g1_n2048")
    f = run_direct(g["y"], g["mass"], options=(("direct_symmetric", 1), ("direct_sym_tile", 512)))
    assert rel_err_per_body(f, g["f_openmp"], 2048) <= TOL64
    assert rel_err_per_body(f, g["f_block"], 2048) <= TOL64


def test_direct_symmetric_is_the_large_n_default(oracle64):
    """N = 262,144 takes the symmetric path automatically (tile 2048): sampled check + third-law property."""
    n = 262144
    y, m = universe(n)
    from nbody_b200 import Engine
    with Engine() as e:
        assert e.init(y, m)
        fb = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), fb)
        assert e.last_direct_path() == 2048
    f = run_direct(y, m).reshape(6, n)
    t = np.unique(np.concatenate([[0, n // 2, n - 1], np.random.RandomState(9).randint(0, n, 125)]))
    ref = oracle64.accel_subset(y, m, t)
    assert rel_err_per_body(f[3:, t], ref, t.size) <= TOL64
    force = (f[3:] * m[None, :]).sum(axis=1)
    scale = np.abs(f[3:] * m[None, :]).sum(axis=1)
    assert np.all(np.abs(force) <= 1e-11 * scale)


@pytest.mark.parametrize("precision,n,edge", [("f64", 8192, 256), ("f64", 16384, 256), ("f64", 65536, 512), ("f32", 16384, 256),
                                              ("f32", 65536, 512)])
def test_direct_symmetric_small_tiles_are_automatic(oracle64, precision, n, edge):
    """From N = 8,192 (FP32: 16,384) the symmetric path is the default: tiles of 256 / 512 bodies run CTAs of 2 / 4 warps
    (FP32: 1 / 2), several per SM. Sampled against the oracle, and Newton's third law over all bodies."""
    y, m = universe(n, precision)
    from nbody_b200 import Engine
    with Engine(precision=precision) as e:
        assert e.init(y, m)
        fb = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), fb)
        assert e.last_direct_path() == edge
        f = e.read_buffer(fb).astype(np.float64).reshape(6, n)
    t = np.unique(np.concatenate([[0, n // 2, n - 1], np.random.RandomState(3).randint(0, n, 125)]))
    ref = oracle64.accel_subset(y.astype(np.float64), m.astype(np.float64), t)
    assert rel_err_per_body(f[3:, t], ref, t.size) <= (TOL64 if precision == "f64" else TOL32)
    force = (f[3:] * m[None, :]).sum(axis=1)
    scale = np.abs(f[3:] * m[None, :]).sum(axis=1)
    assert np.all(np.abs(force) <= (1e-11 if precision == "f64" else 1e-3) * scale)


@pytest.mark.parametrize("n,tile,shape", [(4096, 1024, 2), (5000, 512, 1), (4096, 1024, 4), (5000, 512, 5), (3000, 1024, 4),
                                          (2048, 256, 4), (3000, 512, 4), (8192, 2048, 4)])
def test_direct_symmetric_tiles_fp32(oracle32, oracle64, n, tile, shape):
    """Shapes 4 and 5 are the packed fma.rn.f32x2 kernels (two column bodies per instruction)."""
    rng = np.random.RandomState(n)
    y = rng.uniform(-50, 50, 6 * n).astype(np.float32)
    m = rng.uniform(0.1, 2.0, n).astype(np.float32)
    y[1] = y[0]
    y[n + 1] = y[n]
    y[2 * n + 1] = y[2 * n]            # bodies 0 and 1 coincide: clamped pair, exactly zero contribution
    f = run_direct(y, m, precision="f32", options=(("direct_symmetric", 1), ("direct_sym_tile", tile), ("direct_sym_shape", shape)))
    assert np.all(np.isfinite(f))
    ref32 = oracle32.fcompute_openmp(y, m)
    truth = oracle64.fcompute_openmp(y.astype(np.float64), m.astype(np.float64))
    assert np.array_equal(f[:3 * n], y[3 * n:])
    assert rel_err_per_body(f, ref32, n) <= TOL32
    assert rel_err_per_body(f, truth, n) <= max(2 * rel_err_per_body(ref32, truth, n), 2e-6)


@pytest.mark.parametrize("shape", [0, 1])
def test_direct_symmetric_tiles_redo_a_tile_with_the_clamp(oracle64, shape):
    """The symmetric tiles' first pass leaves out max(r^2, MinDistance) and watches for pairs closer than 1e-4; a tile
    that holds one is redone with the clamp (diagonal tiles clamp from the start: every body meets itself there).
    Bodies placed 1e-6 apart and exactly on top of each other, in the SAME tile and in DIFFERENT tiles: the result is
    the reference's (nbody_data::force, nbody_data.cpp:39-42) for them and for everybody else."""
    n, tile = 8192, 1024
    rng = np.random.RandomState(7)
    pos = rng.uniform(-50, 50, (3, n))
    vel = rng.uniform(-1, 1, (3, n))
    m = rng.uniform(0.1, 1.0, n)
    close = [(10, 5000, 1e-6), (20, 7000, 0.0), (3000, 3001, 1e-6), (4100, 4200, 0.0), (100, 8191, 3e-5)]
    for i, j, gap in close:
        pos[:, j] = pos[:, i]
        pos[0, j] += gap
    y = np.concatenate([pos.reshape(-1), vel.reshape(-1)])
    want = oracle64.fcompute_openmp(y, m)
    f = run_direct(y, m, options=(("direct_symmetric", 1), ("direct_sym_tile", tile), ("direct_sym_shape", shape)))
    assert np.isfinite(f).all()
    assert rel_err_per_body(f, want, n) <= TOL64
    g = run_direct(y, m, options=(("direct_symmetric", 0),))          # ordered-pair kernel: same answer
    assert rel_err_per_body(f, g, n) <= TOL64


def test_use_nccl_needs_distinct_devices_and_otherwise_changes_nothing(oracle64):
    """use_nccl=1 (the reference's factory parameter): one lane has nothing to exchange; a device list that repeats a
    device cannot form NCCL communicators -- the option is refused, the peer path stays, results are unchanged."""
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n2048")
    want = oracle64.fcompute_openmp(g["y"], g["mass"])
    with Engine(devices="0") as e:
        assert e.set_option("use_nccl", 1) == 0
        assert e.init(g["y"], g["mass"])
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), f)
        assert rel_err_per_body(e.read_buffer(f), want, 2048) <= TOL64
    with Engine(devices="0,0") as e:
        assert e.set_option("use_nccl", 1) != 0 and "repeats device" in e.last_error()
        assert e.init(g["y"], g["mass"])
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), f)
        assert rel_err_per_body(e.read_buffer(f), want, 2048) <= TOL64
        assert "peer" in e.print_info()
