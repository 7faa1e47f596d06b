"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/liboracle_f{64,32}.so
(the plain-C restatement in nbody_oracle.c). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(precision="f64"):
    return os.path.join(_HERE, "liboracle_%s.so" % precision)


def build(force=False):
    """gcc the restatement (and, where /root/reference is mounted, the reference itself)."""
    targets = ["oracle"]
    if os.path.isdir("/root/reference/nbody"):
        targets += ["ref", "ref_v4", "ref_cuda"]
    cmd = ["make", "-C", _HERE, "-j8"] + (["-B"] if force else []) + targets
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


def build_sim():
    """nbody-simulation in short form with both overlay patches (oracle/_ref/nbody_sim_f{64,32}); links libnb200, so it
    is built after the kernel libraries. Only where /root/reference is mounted."""
    if os.path.isdir("/root/reference/nbody"):
        subprocess.run(["make", "-C", _HERE, "-j8", "sim"], check=True, stdout=subprocess.DEVNULL)


def sim_path(precision="f64"):
    return os.path.join(_HERE, "_ref", "nbody_sim_%s" % precision)


def load(precision="f64"):
    if precision in _LIBS:
        return _LIBS[precision]
    if not os.path.exists(lib_path(precision)):
        build()
    lib = C.CDLL(lib_path(precision))
    real = C.c_double if precision == "f64" else C.c_float
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int

    def sig(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("orc_real_size", i32)
    sig("orc_max_threads", i32)
    sig("orc_set_threads", None, i32)
    sig("orc_fcompute_openmp", None, sz, vp, vp, vp)
    sig("orc_fcompute_block", None, sz, vp, vp, vp)
    sig("orc_accel_subset", None, sz, vp, vp, vp, sz, vp)
    sig("orc_accel_subset_ld", None, sz, vp, vp, vp, sz, vp)
    sig("orc_fmadd_inplace", None, vp, vp, real, sz)
    sig("orc_fmadd", None, vp, vp, vp, real, sz)
    sig("orc_fmaddn_inplace", None, vp, vp, vp, sz, sz)
    sig("orc_fmaddn", None, vp, vp, vp, vp, sz, sz)
    sig("orc_fmaddn_corr", None, vp, vp, vp, vp, sz, sz)
    sig("orc_fmaxabs", real, vp, sz)
    sig("orc_clamp", None, vp, real, sz)
    for name in ("left", "right", "parent", "next_down", "skip"):
        sig("orc_heap_" + name, sz, sz)
    sig("orc_heap_next_up", sz, sz, sz)
    sig("orc_heap_build", i32, sz, vp, vp, real, vp, vp, vp, vp, vp)
    sig("orc_heap_rebuild", None, sz, vp, real, vp, vp, vp, vp, vp)
    sig("orc_fcompute_bh", None, sz, vp, vp, vp, vp, vp, i32, vp, vp, vp)
    sig("orc_bh_subset", None, sz, vp, vp, vp, sz, vp)
    sig("orc_run_euler", None, sz, vp, vp, real, real)
    sig("orc_run_rk4", None, sz, vp, vp, real, real)
    sig("orc_statistics", None, sz, vp, vp, i32, vp)
    lib.dtype = np.dtype(np.float64 if precision == "f64" else np.float32)
    assert lib.orc_real_size() == lib.dtype.itemsize
    _LIBS[precision] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, precision="f64"):
        self.lib = load(precision)
        self.dtype = self.lib.dtype

    def _a(self, x):
        return np.ascontiguousarray(x, dtype=self.dtype)

    def threads(self):
        return self.lib.orc_max_threads()

    # direct
    def fcompute_openmp(self, y, mass):
        y, mass = self._a(y), self._a(mass)
        f = np.empty_like(y)
        self.lib.orc_fcompute_openmp(mass.size, _p(y), _p(mass), _p(f))
        return f

    def fcompute_block(self, y, mass):
        y, mass = self._a(y), self._a(mass)
        assert mass.size % 64 == 0
        f = np.empty_like(y)
        self.lib.orc_fcompute_block(mass.size, _p(y), _p(mass), _p(f))
        return f

    def accel_subset(self, y, mass, targets, long_double=False):
        y, mass = self._a(y), self._a(mass)
        t = np.ascontiguousarray(targets, dtype=np.uint64)
        acc = np.empty(3 * t.size, dtype=self.dtype)
        fn = self.lib.orc_accel_subset_ld if long_double else self.lib.orc_accel_subset
        fn(mass.size, _p(y), _p(mass), _p(t), t.size, _p(acc))
        return acc.reshape(3, t.size)

    # state ops (operate on copies, return results)
    def fmadd_inplace(self, a, b, c):
        a = self._a(a).copy()
        self.lib.orc_fmadd_inplace(_p(a), _p(self._a(b)), c, a.size)
        return a

    def fmadd(self, b, c, d):
        b, c = self._a(b), self._a(c)
        a = np.empty_like(b)
        self.lib.orc_fmadd(_p(a), _p(b), _p(c), d, a.size)
        return a

    def _ptrs(self, arrays):
        keep = [self._a(x) for x in arrays]
        arr = (C.c_void_p * max(1, len(keep)))(*[x.ctypes.data for x in keep])
        return keep, arr

    def fmaddn_inplace(self, a, bs, c):
        a = self._a(a).copy()
        keep, arr = self._ptrs(bs)
        c = self._a(c)
        self.lib.orc_fmaddn_inplace(_p(a), arr, _p(c), c.size, a.size)
        return a

    def fmaddn(self, a, b, cs, d):
        a = self._a(a).copy()
        keep, arr = self._ptrs(cs)
        d = self._a(d)
        bb = None if b is None else self._a(b)
        self.lib.orc_fmaddn(_p(a), _p(bb), arr, _p(d), d.size, a.size)
        return a

    def fmaddn_corr(self, a, corr, bs, c):
        a, corr = self._a(a).copy(), self._a(corr).copy()
        keep, arr = self._ptrs(bs)
        c = self._a(c)
        self.lib.orc_fmaddn_corr(_p(a), _p(corr), arr, _p(c), c.size, a.size)
        return a, corr

    def fmaxabs(self, a):
        a = self._a(a)
        return self.dtype.type(self.lib.orc_fmaxabs(_p(a), a.size))

    def clamp(self, y, b):
        y = self._a(y).copy()
        self.lib.orc_clamp(_p(y), b, y.size // 6)
        return y

    # heap
    def heap_build(self, y, mass, ratio):
        y, mass = self._a(y), self._a(mass)
        n = mass.size
        t = dict(n=n, ratio=ratio,
                 xyzr=np.zeros((2 * n, 4), dtype=self.dtype), mass=np.zeros(2 * n, dtype=self.dtype),
                 bmin=np.zeros((2 * n, 3), dtype=self.dtype), bmax=np.zeros((2 * n, 3), dtype=self.dtype),
                 body_n=np.zeros(2 * n, dtype=np.int64))
        rc = self.lib.orc_heap_build(n, _p(y), _p(mass), ratio, _p(t["xyzr"]), _p(t["mass"]), _p(t["bmin"]),
                                     _p(t["bmax"]), _p(t["body_n"]))
        if rc != 0:
            raise ValueError("kd-heap needs N = 2^k")
        return t

    def heap_rebuild(self, t, y):
        y = self._a(y)
        self.lib.orc_heap_rebuild(t["n"], _p(y), t["ratio"], _p(t["xyzr"]), _p(t["mass"]), _p(t["bmin"]),
                                  _p(t["bmax"]), _p(t["body_n"]))
        return t

    def fcompute_bh(self, y, mass, tree, stackless=True):
        y, mass = self._a(y), self._a(mass)
        f = np.zeros_like(y)
        v, k = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.orc_fcompute_bh(mass.size, _p(y), _p(mass), _p(tree["xyzr"]), _p(tree["mass"]), _p(tree["body_n"]),
                                 1 if stackless else 0, _p(f), C.byref(v), C.byref(k))
        return f, v.value, k.value

    def bh_subset(self, tree, leaves):
        """Accelerations (3 x nt) of the targets at the given leaf positions, walking `tree` like simple_bh."""
        t = np.ascontiguousarray(leaves, dtype=np.uint64)
        acc = np.empty(3 * t.size, dtype=self.dtype)
        self.lib.orc_bh_subset(tree["n"], _p(tree["xyzr"]), _p(tree["mass"]), _p(t), t.size, _p(acc))
        return acc.reshape(3, t.size)

    # solvers / statistics
    def run(self, solver, y, mass, dt, max_time):
        y, mass = self._a(y).copy(), self._a(mass)
        fn = {"euler": self.lib.orc_run_euler, "rk4": self.lib.orc_run_rk4}[solver]
        fn(mass.size, _p(y), _p(mass), dt, max_time)
        return y

    def statistics(self, y, mass, with_energy=True):
        y, mass = self._a(y), self._a(mass)
        out = np.zeros(11, dtype=self.dtype)
        self.lib.orc_statistics(mass.size, _p(y), _p(mass), 1 if with_energy else 0, _p(out))
        return dict(P=out[0:3], L=out[3:6], Ekin=out[6], Epot=out[7], C=out[8:11])


def load_table(path):
    """Reader for the reference's G1 text tables (nbody_data::load, nbody_data.cpp:442-512):
    whitespace columns X Y Z Vx Vy Vz Mass [Radius ...], '//' starts a comment.
    Returns (y[6N], mass[N]) in float64."""
    rows = []
    with open(path) as fh:
        for line in fh:
            cut = line.find("//")
            if cut >= 0:
                line = line[:cut]
            cols = line.split()
            if not cols:
                continue
            if len(cols) < 7:
                raise ValueError("need >= 7 columns: %r" % line)
            rows.append([float(c) for c in cols[:7]])
    a = np.array(rows, dtype=np.float64)
    y = np.concatenate([a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4], a[:, 5]])
    return y, a[:, 6].copy()
