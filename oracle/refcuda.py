"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- ctypes view of oracle/_ref/libnbref_cuda_f{64,32}.so: the reference's own
CUDA kernels (nbody_engine_cuda_impl.cu) recompiled for sm_100a, driven by oracle/ref_cuda_harness.cu."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(precision="f64"):
    return os.path.join(_HERE, "_ref", "libnbref_cuda_%s.so" % precision)


def available(precision="f64"):
    return os.path.exists(lib_path(precision))


def load(precision="f64"):
    if precision not in _LIBS:
        lib = C.CDLL(lib_path(precision), mode=C.RTLD_LOCAL)
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        lib.nbrefcu_direct.restype = i32
        lib.nbrefcu_direct.argtypes = [vp, vp, sz, i32, i32, vp, C.POINTER(C.c_double)]
        lib.nbrefcu_bh_stackless.restype = i32
        lib.nbrefcu_bh_stackless.argtypes = [vp, sz, vp, vp, vp, i32, i32, vp, C.POINTER(C.c_double)]
        lib.dtype = np.dtype(np.float64 if precision == "f64" else np.float32)
        _LIBS[precision] = lib
    return _LIBS[precision]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def direct(y, mass, block_size=64, reps=1, precision="f64"):
    """(f, ms): kfcompute + kfcompute_xyz; N must be a multiple of block_size (the kernel has no bounds checks)."""
    lib = load(precision)
    y = np.ascontiguousarray(y, dtype=lib.dtype)
    mass = np.ascontiguousarray(mass, dtype=lib.dtype)
    assert mass.size % block_size == 0
    f = np.empty_like(y)
    ms = C.c_double(0)
    rc = lib.nbrefcu_direct(_p(y), _p(mass), mass.size, block_size, reps, _p(f), C.byref(ms))
    if rc != 0:
        raise RuntimeError("reference CUDA kernel failed (%d)" % rc)
    return f, ms.value


def bh_stackless(y, xyzr, node_mass, body_n, block_size=256, reps=1, precision="f64"):
    """(f, ms): kfcompute_heap_bh_stackless on a caller-supplied tree (2N nodes)."""
    lib = load(precision)
    y = np.ascontiguousarray(y, dtype=lib.dtype)
    xyzr = np.ascontiguousarray(xyzr, dtype=lib.dtype)
    node_mass = np.ascontiguousarray(node_mass, dtype=lib.dtype)
    body_n = np.ascontiguousarray(body_n, dtype=np.int32)
    n = node_mass.size // 2
    assert n % block_size == 0
    f = np.empty_like(y)
    ms = C.c_double(0)
    rc = lib.nbrefcu_bh_stackless(_p(y), n, _p(xyzr), _p(node_mass), _p(body_n), block_size, reps, _p(f), C.byref(ms))
    if rc != 0:
        raise RuntimeError("reference CUDA BH kernel failed (%d)" % rc)
    return f, ms.value
