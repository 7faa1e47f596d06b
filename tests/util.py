"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

_CACHE = {}


def gpu_available():
    try:
        from nbody_b200 import device_count
        return device_count("f64") > 0
    except Exception:
        return False


def numpy_universe(n, seed=5489):
    """Two-disk galaxy collision with the geometry of nbody_data::make_universe
    (nbody_data.cpp:263-324); numpy RNG, so NOT the reference's libstdc++ stream.
    Used only where the compiled reference (oracle/_ref) is not available."""
    rng = np.random.RandomState(seed % (2 ** 32))
    half = n // 2
    radius, gm = 50.0, 1000.0
    bh = gm * 0.999
    star = (gm - bh) / (half - 1)
    vgal = np.sqrt(gm * gm / radius ** 2 * radius / (2 * gm)) / 3.0
    pos, vel, mass = [], [], []
    for sign in (-1.0, +1.0):
        center = np.array([50.0 + sign * radius, 50.0, 50.0])
        gv = np.array([0.0, -sign * vgal, 0.0])
        pts = []
        while len(pts) < half - 1:
            c = rng.uniform(-radius, radius, size=(half, 3))
            c = c[np.sqrt((c ** 2).sum(axis=1)) <= radius]
            pts.extend(c.tolist())
        r = np.array(pts[:half - 1])
        rlen = np.sqrt((r ** 2).sum(axis=1))
        r[:, 2] *= 0.3
        d = r
        v = np.cross(d, np.array([0.0, 0.0, 1.0]))
        v /= np.maximum(np.sqrt((v ** 2).sum(axis=1))[:, None], 1e-30)
        meff = (rlen / radius) ** 3 * (half - 1) * star + bh
        dist = np.sqrt((d ** 2).sum(axis=1))
        v *= np.sqrt(meff / dist)[:, None]
        pos.append(center[None, :])
        pos.append(r + center)
        vel.append(gv[None, :])
        vel.append(v + gv)
        mass.append(np.array([bh]))
        mass.append(np.full(half - 1, star))
    p, v, m = np.concatenate(pos), np.concatenate(vel), np.concatenate(mass)
    y = np.concatenate([p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2]])
    return y, m


def universe(n, precision="f64"):
    """G1 synthetic galaxy pair with N bodies: from the compiled reference when present."""
    key = (n, precision)
    if key in _CACHE:
        return _CACHE[key]
    from oracle import refharness as R
    dtype = np.float64 if precision == "f64" else np.float32
    if R.available(precision):
        lib = R.load(precision)
        d = R.Data(lib).make_universe(n // 2)
        y, m = d.export()
        d.close()
        assert m.size == n
    else:
        y, m = numpy_universe(n)
        y, m = y.astype(dtype), m.astype(dtype)
    _CACHE[key] = (y, m)
    return y, m
