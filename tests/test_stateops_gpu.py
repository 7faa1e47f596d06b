"""Engine conformance on the GPU, mirroring test/engine/test_nbody_engine.cpp:23-470,560-617,761-1011
through the Python host mirror (which calls the C ABI): buffers, fmadd*, fmaddn*, fmaddn_corr, fmaxabs,
clamp and the negative branches (log and return, never crash). Results are checked against the CPU
oracle, bit-exactly where the reference's test uses eps = DBL_EPSILON on integer data."""
import numpy as np
import pytest

from conftest import load_golden_npz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[("f64", "0"), ("f32", "0"), ("f64", "0,0")], ids=["f64", "f32", "f64-2lanes"])
def eng(request):
    from nbody_b200 import Engine
    precision, devices = request.param
    g = load_golden_npz("g1_n128", precision)
    e = Engine(precision=precision, devices=devices)
    assert e.init(g["y"], g["mass"])
    yield e
    e.close()


def oracle_for(e):
    from oracle.oracle import Oracle
    return Oracle("f64" if e.dtype == np.float64 else "f32")


def test_mem_and_memcpy(eng):
    m = eng.create_buffer(1024)
    assert m is not None and m.size() == 1024
    eng.free_buffer(m)
    data = np.arange(8, dtype=eng.dtype)
    mem = eng.create_buffer(data.nbytes)
    sub = eng.create_buffer(data.nbytes // 2)
    eng.write_buffer(mem, data)
    assert np.array_equal(eng.read_buffer(mem), data)
    eng.free_buffer(mem)
    eng.free_buffer(sub)
    eng.free_buffer(None)          # free_buffer(nullptr) is a no-op (euler's dtor relies on it)


def test_copy_buffer(eng):
    n = eng.problem_size()
    d1 = np.arange(n, dtype=eng.dtype)
    m1, m2 = eng.create_buffer(d1.nbytes), eng.create_buffer(d1.nbytes)
    eng.write_buffer(m1, d1)
    eng.copy_buffer(m2, m1)
    assert np.array_equal(eng.read_buffer(m2), d1)
    eng.free_buffer(m1)
    eng.free_buffer(m2)


def test_fill_buffer_33_elements(eng):
    m = eng.create_buffer(33 * eng.dtype.itemsize)
    eng.fill_buffer(m, 777)
    assert np.all(eng.read_buffer(m) == 777)
    eng.free_buffer(m)
    z = eng.create_buffer(0)
    eng.fill_buffer(z, 1)
    assert eng.read_buffer(z).size == 0
    eng.free_buffer(z)


def _ints(rng, n, dtype, hi=10000):
    # FP32 holds integers exactly only below 2^24: keep products small there
    return rng.randint(0, hi if dtype == np.float64 else 64, n).astype(dtype)


def test_fmadd_inplace_and_fmadd(eng):
    rng = np.random.RandomState(11)
    n = eng.problem_size()
    a, b, c = (_ints(rng, n, eng.dtype) for _ in range(3))
    ma, mb, mc = eng.create_buffer(a.nbytes), eng.create_buffer(a.nbytes), eng.create_buffer(a.nbytes)
    eng.write_buffer(ma, a)
    eng.write_buffer(mb, b)
    eng.write_buffer(mc, c)
    eng.fmadd_inplace(ma, mb, 5)
    assert np.array_equal(eng.read_buffer(ma), a + 5 * b)
    eng.fmadd(ma, mb, mc, 5)
    assert np.array_equal(eng.read_buffer(ma), b + c * 5)
    eng.fmadd(mb, mb, mc, 2)                    # a aliases b, as nbody_engine::fmaddn does
    assert np.array_equal(eng.read_buffer(mb), b + c * 2)
    eng.fmadd_inplace(mc, mc, 0.5)              # y += y*c, used by test_fcompute (:548-549)
    assert np.array_equal(eng.read_buffer(mc), c + c * eng.dtype.type(0.5))
    for m in (ma, mb, mc):
        eng.free_buffer(m)


@pytest.mark.parametrize("csize", [1, 3, 7, 35, 60])
def test_fmaddn_family_vs_oracle(eng, csize):
    """fmaddn_inplace / fmaddn / fmaddn(NULL b) with zero coefficients injected at k = 0 and k = csize/2
    (test_nbody_engine.cpp:188-375); 35 = rkfeagin14 stage count, 60 > one fused launch (48 terms)."""
    o = oracle_for(eng)
    rng = np.random.RandomState(csize)
    n = eng.problem_size()
    a, b = rng.rand(n).astype(eng.dtype), rng.rand(n).astype(eng.dtype)
    ks = [rng.rand(n).astype(eng.dtype) for _ in range(csize)]
    cf = rng.uniform(-2, 2, csize).astype(eng.dtype)
    cf[0] = 0
    cf[csize // 2] = 0
    ma, mb = eng.create_buffer(a.nbytes), eng.create_buffer(a.nbytes)
    mk = eng.create_buffers(a.nbytes, csize)
    assert len(mk) == csize
    eng.write_buffer(mb, b)
    for m, k in zip(mk, ks):
        eng.write_buffer(m, k)
    eng.write_buffer(ma, a)
    eng.fmaddn_inplace(ma, mk, cf)
    assert np.array_equal(eng.read_buffer(ma), o.fmaddn_inplace(a, ks, cf))
    eng.write_buffer(ma, a)
    eng.fmaddn(ma, mb, mk, cf)
    assert np.array_equal(eng.read_buffer(ma), o.fmaddn(a, b, ks, cf))
    eng.write_buffer(ma, a)
    eng.fmaddn(ma, None, mk, cf)
    assert np.array_equal(eng.read_buffer(ma), o.fmaddn(a, None, ks, cf))
    # all-zero coefficients: with b the reference leaves a untouched, without b it zero-fills (nbody_engine.cpp:87-112)
    z = np.zeros(csize, dtype=eng.dtype)
    eng.write_buffer(ma, a)
    eng.fmaddn(ma, mb, mk, z)
    assert np.array_equal(eng.read_buffer(ma), a)
    eng.fmaddn(ma, None, mk, z)
    assert np.all(eng.read_buffer(ma) == 0)
    # fmaddn(a, a, ...) -- in-place through the b argument
    eng.write_buffer(ma, a)
    eng.fmaddn(ma, ma, mk, cf)
    assert np.array_equal(eng.read_buffer(ma), o.fmaddn(a, a, ks, cf))
    eng.free_buffer(ma)
    eng.free_buffer(mb)
    eng.free_buffers(mk)


def test_fmaddn_corr_is_compensated(eng):
    """test_fmaddn_corr (:377-437): 1 + 10 * eps/2 is only reached with Kahan summation."""
    eps = np.finfo(eng.dtype).eps
    n = eng.problem_size()
    one = np.ones(n, dtype=eng.dtype)
    ma, mc = eng.create_buffer(one.nbytes), eng.create_buffer(one.nbytes)
    mb = eng.create_buffers(one.nbytes, 10)
    eng.write_buffer(ma, one)
    eng.fill_buffer(mc, 0)
    for m in mb:
        eng.write_buffer(m, one)
    c = np.full(10, eps / 2, dtype=eng.dtype)
    eng.fmaddn_corr(ma, mc, mb, c)
    res = eng.read_buffer(ma)
    assert np.abs((one + eng.dtype.type(10) * (eps / 2)) - res).max() <= eps
    eng.write_buffer(ma, one)
    eng.fmaddn_inplace(ma, mb, c)
    assert np.all(eng.read_buffer(ma) == 1)      # uncompensated: every term is lost
    eng.free_buffer(ma)
    eng.free_buffer(mc)
    eng.free_buffers(mb)


def test_fmaddn_corr_vs_oracle_bit_exact(eng):
    o = oracle_for(eng)
    rng = np.random.RandomState(5)
    n = eng.problem_size()
    a, corr = rng.rand(n).astype(eng.dtype), (rng.rand(n) * 1e-17).astype(eng.dtype)
    ks = [rng.rand(n).astype(eng.dtype) for _ in range(7)]
    cf = np.array([1e-3, 0, 2.5, -1, 0, 1e-9, 7], dtype=eng.dtype)
    ma, mc = eng.create_buffer(a.nbytes), eng.create_buffer(a.nbytes)
    mk = eng.create_buffers(a.nbytes, 7)
    eng.write_buffer(ma, a)
    eng.write_buffer(mc, corr)
    for m, k in zip(mk, ks):
        eng.write_buffer(m, k)
    eng.fmaddn_corr(ma, mc, mk, cf)
    oa, oc = o.fmaddn_corr(a, corr, ks, cf)
    assert np.array_equal(eng.read_buffer(ma), oa)
    assert np.array_equal(eng.read_buffer(mc), oc)
    eng.free_buffer(ma)
    eng.free_buffer(mc)
    eng.free_buffers(mk)


def test_fmaxabs(eng):
    rng = np.random.RandomState(3)
    n = eng.problem_size()
    a = (rng.randint(0, 10000, n) - 9000).astype(eng.dtype)
    m = eng.create_buffer(a.nbytes)
    eng.write_buffer(m, a)
    assert eng.fmaxabs(m, default=2878767678687.0) == np.abs(a).max()
    a[:] = 0
    a[n - 1] = -3.5                                 # maximum in the very last element (tail handling)
    eng.write_buffer(m, a)
    assert eng.fmaxabs(m) == 3.5
    eng.free_buffer(m)
    odd = eng.create_buffer(33 * eng.dtype.itemsize)
    eng.fill_buffer(odd, -2)
    assert eng.fmaxabs(odd) == 2
    eng.free_buffer(odd)
    empty = eng.create_buffer(0)
    assert eng.fmaxabs(empty, default=123.0) == 0   # size 0 -> 0 (nbody_engine_cuda.cpp:511-515)
    eng.free_buffer(empty)


def test_clamp(eng):
    """test_clamp (:560-617)."""
    n = eng.problem_size()
    a = np.where(np.arange(n) & 1, 4, -4).astype(eng.dtype)
    m = eng.create_buffer(a.nbytes)
    eng.write_buffer(m, a)
    eng.clamp(m, 1)
    b = eng.read_buffer(m)
    exp = a.copy()
    exp[:n // 2] = np.where(np.arange(n // 2) & 1, 2, -2)
    assert np.array_equal(b, exp)
    eng.free_buffer(m)


def test_time_step_and_counters(eng):
    eng.set_time(0.0)
    eng.set_step(0)
    eng.advise_time(0.25)
    assert eng.get_time() == 0.25 and eng.get_step() == 1
    eng.set_step(7)
    assert eng.get_step() == 7
    assert eng.problem_size() == 6 * 128 and eng.get_y().size() == 6 * 128 * eng.dtype.itemsize


def test_negative_branches(eng):
    """test_negative_branches (:761-1011): foreign memory, NULL, csize > b.size(), byte-sized buffers."""
    from nbody_b200 import Engine, Memory

    class Fake:
        def size(self):
            return 0

    fake = Fake()
    ps = eng.problem_size()
    y = eng.create_buffer(ps)                     # problem_size() BYTES, as the reference's test does
    before = eng.launch_count()
    eng.fcompute(0, fake, y)
    eng.fcompute(0, y, fake)
    eng.fcompute(0, y, y)
    assert eng.read_buffer(fake) is None
    eng.write_buffer(fake, None)
    eng.write_buffer(y, None)
    eng.copy_buffer(fake, y)
    eng.copy_buffer(y, fake)
    eng.fmadd_inplace(fake, y, 1)
    eng.fmadd_inplace(y, fake, 1)
    eng.fmadd(fake, y, y, 0)
    eng.fmadd(y, y, fake, 0)
    eng.fmadd(y, fake, y, 0)
    one = eng.create_buffers(ps, 1)
    zero, unit = np.zeros(1, dtype=eng.dtype), np.ones(1, dtype=eng.dtype)
    eng.fmaddn_inplace(fake, one, zero)
    eng.fmaddn_inplace(y, one, None, 1)
    eng.fmaddn_inplace(y, [fake], zero)
    eng.fmaddn_inplace(y, one, zero, 100)
    eng.fmaddn_inplace(y, [None], unit)
    eng.fmaddn_corr(None, y, one, zero)
    eng.fmaddn_corr(y, None, one, zero)
    eng.fmaddn_corr(y, one[0], [None], zero)
    eng.fmaddn_corr(y, one[0], one, None, 1)
    eng.fmaddn_corr(y, one[0], one, zero, 100)
    eng.fmaddn_corr(y, one[0], [None], unit)
    eng.fmaddn(fake, y, one, zero, 200000)
    eng.fmaddn(fake, y, one, zero)
    eng.fmaddn(fake, None, one, zero)
    eng.fmaddn(y, y, one, None, 1)
    eng.fmaddn(y, None, one, None, 1)
    eng.fmaddn(y, fake, one, zero)
    eng.fmaddn(y, y, [None], unit)
    assert eng.fmaxabs(fake, default=0.0) == 0.0
    # a buffer of another engine is foreign too
    other = Engine(precision="f64" if eng.dtype == np.float64 else "f32")
    ob = other.create_buffer(ps)
    eng.fmadd_inplace(y, ob, 1)
    eng.copy_buffer(y, ob)
    other.free_buffer(ob)
    other.close()
    # size mismatch
    big = eng.create_buffer(2 * ps)
    eng.copy_buffer(big, y)
    eng.fmadd_inplace(big, y, 1)
    eng.free_buffer(big)
    # every call above was rejected before any kernel launch
    assert eng.launch_count() == before
    eng.free_buffers(one)
    eng.free_buffer(y)
    assert isinstance(eng.get_y(), Memory)


def test_invalid_device_strings():
    """Factory must refuse "", "a", "0,a", "-1", "9999" (test_nbody_engine.cpp:1227-1240)."""
    from nbody_b200 import Engine
    for bad in ("", "a", "0,a", "-1", "9999"):
        with pytest.raises(ValueError):
            Engine(devices=bad)
    with pytest.raises(ValueError):
        Engine(kind="bh", tree_layout="tree")


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_fmaxabs_nan_follows_the_reference_loop(ref64, devices):
    """fmaxabs of the reference: result = fabs(a[0]); then `if(v > result)` over all elements
    (nbody_engine_openmp.cpp:284-296). A NaN in element 0 therefore IS the result, a NaN anywhere else is skipped --
    the engine does the same on every shard layout, so a run that blew up branches as it does on the CPU engines."""
    from nbody_b200 import Engine
    from oracle import refharness as R
    g = load_golden_npz("g1_n128")
    n6 = 6 * 128
    d = R.Data(ref64).import_(g["y"], g["mass"])
    cpu = R.Engine(ref64, engine="simple")
    assert cpu.init(d)
    with Engine(devices=devices) as e:
        assert e.init(g["y"], g["mass"])
        buf = e.create_buffer(n6 * 8)
        for where in (None, 0, 1, n6 // 2, n6 - 1):
            a = np.linspace(-3.0, 2.0, n6)
            if where is not None:
                a[where] = np.nan
            e.write_buffer(buf, a)
            got = e.fmaxabs(buf)
            m = cpu.new_buffer(a)
            want = cpu.fmaxabs(m)
            cpu.free_buffer(m)
            assert (np.isnan(got) and np.isnan(want)) or got == want, (where, got, want)
            assert np.isnan(got) == (where == 0)
    cpu.close()
    d.close()


@pytest.mark.parametrize("devices", ["0", "0,0", "0,0,0,0"])
def test_state_sized_buffer_created_before_the_bodies_are_known(devices):
    """A buffer of 6N reals allocated before set_bodies / init (the adapter allows create_buffer before init) is a
    state vector afterwards: with one shard it already has that layout, with several each lane keeps its own columns
    of what was written into it."""
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n128")
    with Engine(devices=devices) as e:
        early = e.create_buffer(6 * 128 * 8)
        marks = np.arange(6 * 128, dtype=np.float64)
        e.write_buffer(early, marks)
        assert e.init(g["y"], g["mass"])
        assert np.array_equal(e.read_buffer(early), marks)          # contents survive the re-sharding
        e.fcompute(0.0, e.get_y(), early)
        late = e.create_buffer(6 * 128 * 8)
        e.fcompute(0.0, e.get_y(), late)
        assert np.array_equal(e.read_buffer(early), e.read_buffer(late))
        e.fmadd_inplace(early, late, 1.0)
        assert np.array_equal(e.read_buffer(early), 2 * e.read_buffer(late))
