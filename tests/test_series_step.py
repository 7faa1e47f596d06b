"""The one-step series behind the round-2 force kernels (nb200_direct_sym.cuh sym_pair, nb200_bh_group.cuh bhg_force):

    y0  = MUFU.RSQ64H(r2)            a seed of r2^-1/2 with 21 significant bits (only the high word is written)
    y2  = y0 * y0                    exact: 42 bits
    e   = fma(-r2, y2, 1)            = 1 - r2 y0^2, one rounding
    u   = e * fma(e, 15/8, 3/2)
    r^-3 = fma(y0 y2, u, y0 y2)      = y0^3 (1 + 3/2 e + 15/8 e^2);   truncation 35/16 e^3

checked here in numpy against 80-bit arithmetic for seeds anywhere within 2^-20 of the true value (the instruction's
worst case is better than that): the result is r2^-3/2 to a few FP64 ulp, over 60 orders of magnitude of r2."""
import numpy as np


def seeds(r2, rng):
    """A 21-bit seed with a relative error up to 2^-20, low word zero -- what rsqrt.approx.ftz.f64 may return at worst."""
    y = 1.0 / np.sqrt(r2) * (1.0 + rng.uniform(-1, 1, r2.shape) * 2.0 ** -20)
    bits = y.view(np.uint64) & np.uint64(0xFFFFFFFF00000000)
    return bits.view(np.float64)


def test_series_step_gives_r_to_the_minus_three_to_a_few_ulp():
    rng = np.random.RandomState(0)
    r2 = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), 400000)) * np.exp(rng.uniform(-40, 40, 400000) * (rng.rand(400000) < 0.1))
    y0 = seeds(r2, rng)
    y2 = y0 * y0
    assert np.array_equal(y2.astype(np.longdouble), y0.astype(np.longdouble) * y0.astype(np.longdouble))      # exact
    # fma(-r2, y2, 1) with one rounding: evaluate in 80-bit, round once
    e = (np.longdouble(1) - r2.astype(np.longdouble) * y2.astype(np.longdouble)).astype(np.float64)
    u = e * (e * 1.875 + 1.5)
    s3 = y0 * y2
    got = s3 * u + s3
    want = (np.longdouble(1) / (r2.astype(np.longdouble) * np.sqrt(r2.astype(np.longdouble))))
    rel = np.abs((got.astype(np.longdouble) - want) / want).astype(np.float64)
    assert np.abs(e).max() < 2.0 ** -18
    assert rel.max() < 6 * 2.0 ** -53, rel.max()
    # the two-instruction-longer Newton form of round 1 is no more accurate
    h = r2 * y0
    e1 = (np.longdouble(1) - h.astype(np.longdouble) * y0.astype(np.longdouble)).astype(np.float64)
    yv = (y0 * e1) * (e1 * 0.375 + 0.5) + y0
    old = (yv * yv) * yv
    rel_old = np.abs((old.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel.max() <= 1.5 * rel_old.max()
