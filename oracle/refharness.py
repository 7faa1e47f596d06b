"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_ref/libnbref_f{64,32}.so.

The library is the reference's own engines/solvers/data classes compiled
unmodified by oracle/Makefile (see oracle/ref_harness.cpp for the C surface).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(precision="f64", variant=""):
    return os.path.join(_HERE, "_ref", "libnbref_%s%s.so" % (precision, "_" + variant if variant else ""))


def available(precision="f64", variant=""):
    return os.path.exists(lib_path(precision, variant))


def host_variant(precision="f64"):
    """The build of the reference that matches this host best: "v4" (-march=x86-64-v4, AVX-512) where the CPU has
    avx512f/bw/cd/dq/vl and that build exists, else "" (-march=x86-64-v3). The reference itself builds with
    -march=native (pri/vectorize.pri:4); these two fixed levels are what can be prebuilt and shipped."""
    if precision != "f64" or not available(precision, "v4"):
        return ""
    try:
        flags = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
    except OSError:
        return ""
    return "v4" if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags else ""


def _sig(lib, name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


def load(precision="f64", variant=""):
    """Load (once) and type the nbref_* entry points. RTLD_LOCAL: the f64 and f32
    builds define the same C++ symbols with different layouts and must not see
    each other; the nb200 adapter library links its libnbref_* explicitly."""
    key = precision + variant
    if key in _LIBS:
        return _LIBS[key]
    lib = C.CDLL(lib_path(precision, variant), mode=C.RTLD_LOCAL)
    vp, sz, dbl, cs, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_char_p, C.c_int
    _sig(lib, "nbref_coord_size", i32)
    _sig(lib, "nbref_max_threads", i32)
    _sig(lib, "nbref_set_threads", None, i32)
    _sig(lib, "nbref_data_new", vp)
    _sig(lib, "nbref_data_free", None, vp)
    _sig(lib, "nbref_data_make_universe", None, vp, sz, dbl, dbl, dbl)
    _sig(lib, "nbref_data_load_initial", i32, vp, cs, cs)
    _sig(lib, "nbref_data_load", i32, vp, cs)
    _sig(lib, "nbref_data_save", i32, vp, cs)
    _sig(lib, "nbref_data_count", sz, vp)
    _sig(lib, "nbref_data_box_size", sz, vp)
    _sig(lib, "nbref_data_time", dbl, vp)
    _sig(lib, "nbref_data_step", sz, vp)
    _sig(lib, "nbref_data_export", None, vp, vp, vp)
    _sig(lib, "nbref_data_import", None, vp, sz, vp, vp)
    _sig(lib, "nbref_data_is_equal", i32, vp, vp, dbl)
    _sig(lib, "nbref_data_set_check_list", None, vp, cs)
    _sig(lib, "nbref_data_statistics", None, vp, vp, vp)
    _sig(lib, "nbref_engine_create", vp, cs)
    _sig(lib, "nbref_engine_free", None, vp)
    _sig(lib, "nbref_engine_type_name", cs, vp)
    _sig(lib, "nbref_engine_init", i32, vp, vp)
    _sig(lib, "nbref_engine_get_data", None, vp, vp)
    _sig(lib, "nbref_engine_problem_size", sz, vp)
    _sig(lib, "nbref_engine_get_y", vp, vp)
    _sig(lib, "nbref_engine_advise_time", None, vp, dbl)
    _sig(lib, "nbref_engine_get_time", dbl, vp)
    _sig(lib, "nbref_engine_set_time", None, vp, dbl)
    _sig(lib, "nbref_engine_get_step", sz, vp)
    _sig(lib, "nbref_engine_set_step", None, vp, sz)
    _sig(lib, "nbref_engine_compute_count", sz, vp)
    _sig(lib, "nbref_engine_print_info", None, vp)
    _sig(lib, "nbref_engine_fcompute", None, vp, dbl, vp, vp)
    _sig(lib, "nbref_engine_clamp", None, vp, vp, dbl)
    _sig(lib, "nbref_engine_create_buffer", vp, vp, sz)
    _sig(lib, "nbref_engine_free_buffer", None, vp, vp)
    _sig(lib, "nbref_memory_size", sz, vp)
    _sig(lib, "nbref_engine_read_buffer", None, vp, vp, vp)
    _sig(lib, "nbref_engine_write_buffer", None, vp, vp, vp)
    _sig(lib, "nbref_engine_copy_buffer", None, vp, vp, vp)
    _sig(lib, "nbref_engine_fill_buffer", None, vp, vp, dbl)
    _sig(lib, "nbref_engine_fmadd_inplace", None, vp, vp, vp, dbl)
    _sig(lib, "nbref_engine_fmadd", None, vp, vp, vp, vp, dbl)
    _sig(lib, "nbref_engine_fmaddn_inplace", None, vp, vp, vp, sz, vp, sz)
    _sig(lib, "nbref_engine_fmaddn_corr", None, vp, vp, vp, vp, sz, vp, sz)
    _sig(lib, "nbref_engine_fmaddn", None, vp, vp, vp, vp, sz, vp, sz)
    _sig(lib, "nbref_engine_fmaxabs", None, vp, vp, vp)
    _sig(lib, "nbref_foreign_memory_new", vp, sz)
    _sig(lib, "nbref_foreign_memory_free", None, vp)
    _sig(lib, "nbref_engine_time_fcompute", dbl, vp, i32)
    _sig(lib, "nbref_solver_create", vp, cs)
    _sig(lib, "nbref_solver_free", None, vp)
    _sig(lib, "nbref_solver_type_name", cs, vp)
    _sig(lib, "nbref_solver_set_engine", None, vp, vp)
    _sig(lib, "nbref_solver_set_time_step", None, vp, dbl, dbl)
    _sig(lib, "nbref_solver_advise", None, vp, dbl)
    _sig(lib, "nbref_solver_run", i32, vp, vp, dbl, dbl, dbl)
    _sig(lib, "nbref_solver_butcher_check", i32, vp, vp)
    _sig(lib, "nbref_heap_new", vp)
    _sig(lib, "nbref_heap_free", None, vp)
    _sig(lib, "nbref_heap_build", dbl, vp, sz, vp, vp, dbl)
    _sig(lib, "nbref_heap_rebuild", None, vp, sz, vp, dbl)
    _sig(lib, "nbref_heap_export", sz, vp, vp, vp, vp)
    _sig(lib, "nbref_heap_walk_counts", None, vp, sz, sz, sz, vp)
    _sig(lib, "nbref_heap_func", sz, i32, sz, sz)
    lib.precision = precision
    lib.dtype = np.float64 if precision == "f64" else np.float32
    assert lib.nbref_coord_size() == np.dtype(lib.dtype).itemsize
    lib.variant = variant
    _LIBS[key] = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def params(**kw):
    return ";".join("%s=%s" % (k, v) for k, v in kw.items()).encode()


class Data:
    """nbody_data handle."""

    def __init__(self, lib):
        self.lib = lib
        self.h = lib.nbref_data_new()

    def close(self):
        if self.h:
            self.lib.nbref_data_free(self.h)
            self.h = None

    def make_universe(self, stars, box=100.0):
        self.lib.nbref_data_make_universe(self.h, stars, box, box, box)
        return self

    def load(self, path):
        if self.lib.nbref_data_load(self.h, os.fsencode(path)) != 0:
            raise IOError("nbody_data::load failed: %s" % path)
        return self

    def load_initial(self, path, kind):
        if self.lib.nbref_data_load_initial(self.h, os.fsencode(path), kind.encode()) != 0:
            raise IOError("nbody_data::load_initial failed: %s (%s)" % (path, kind))
        return self

    @property
    def count(self):
        return self.lib.nbref_data_count(self.h)

    def export(self):
        n = self.count
        y = np.empty(6 * n, dtype=self.lib.dtype)
        m = np.empty(n, dtype=self.lib.dtype)
        self.lib.nbref_data_export(self.h, _ptr(y), _ptr(m))
        return y, m

    def import_(self, y, mass):
        y = np.ascontiguousarray(y, dtype=self.lib.dtype)
        mass = np.ascontiguousarray(mass, dtype=self.lib.dtype)
        self.lib.nbref_data_import(self.h, mass.size, _ptr(y), _ptr(mass))
        return self

    def is_equal(self, other, eps):
        return bool(self.lib.nbref_data_is_equal(self.h, other.h, eps))

    def statistics(self, engine, check_list="PLVE"):
        out = np.zeros(8)
        self.lib.nbref_data_set_check_list(self.h, check_list.encode())
        self.lib.nbref_data_statistics(self.h, engine.h if engine is not None else None, _ptr(out))
        return dict(dP=out[0], dL=out[1], dE=out[2], dCm=out[3], E=out[4], P=out[5], L=out[6], E0=out[7])


class Engine:
    """Any nbody_engine* (reference CPU engines via the reference factory, or a
    pointer produced by the nb200 adapter library) driven through its virtuals."""

    def __init__(self, lib, handle=None, owned=True, **kw):
        self.lib = lib
        self.h = handle if handle is not None else lib.nbref_engine_create(params(**kw))
        if not self.h:
            raise ValueError("engine factory returned NULL for %r" % (kw,))
        self.owned = owned

    def close(self):
        if self.h and self.owned:
            self.lib.nbref_engine_free(self.h)
        self.h = None

    def type_name(self):
        return self.lib.nbref_engine_type_name(self.h).decode()

    def init(self, data):
        return self.lib.nbref_engine_init(self.h, data.h) == 0

    def get_data(self, data):
        self.lib.nbref_engine_get_data(self.h, data.h)

    def problem_size(self):
        return self.lib.nbref_engine_problem_size(self.h)

    def get_y(self):
        return self.lib.nbref_engine_get_y(self.h)

    def create_buffer(self, nbytes):
        return self.lib.nbref_engine_create_buffer(self.h, nbytes)

    def free_buffer(self, m):
        self.lib.nbref_engine_free_buffer(self.h, m)

    def size(self, m):
        return self.lib.nbref_memory_size(m)

    def write_buffer(self, m, arr):
        arr = np.ascontiguousarray(arr)
        self.lib.nbref_engine_write_buffer(self.h, m, _ptr(arr))

    def read_buffer(self, m, count=None, dtype=None):
        dtype = dtype or self.lib.dtype
        n = count if count is not None else self.size(m) // np.dtype(dtype).itemsize
        out = np.empty(n, dtype=dtype)
        self.lib.nbref_engine_read_buffer(self.h, _ptr(out), m)
        return out

    def new_buffer(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.lib.dtype)
        m = self.create_buffer(arr.nbytes)
        self.write_buffer(m, arr)
        return m

    def copy_buffer(self, a, b):
        self.lib.nbref_engine_copy_buffer(self.h, a, b)

    def fill_buffer(self, a, v):
        self.lib.nbref_engine_fill_buffer(self.h, a, v)

    def fcompute(self, t, y, f):
        self.lib.nbref_engine_fcompute(self.h, t, y, f)

    def clamp(self, y, b):
        self.lib.nbref_engine_clamp(self.h, y, b)

    def fmadd_inplace(self, a, b, c):
        self.lib.nbref_engine_fmadd_inplace(self.h, a, b, c)

    def fmadd(self, a, b, c, d):
        self.lib.nbref_engine_fmadd(self.h, a, b, c, d)

    def _marr(self, ms):
        arr = (C.c_void_p * max(1, len(ms)))(*ms)
        return arr

    def fmaddn_inplace(self, a, b, c, csize=None):
        cc = None if c is None else np.ascontiguousarray(c, dtype=self.lib.dtype)
        n = csize if csize is not None else (0 if cc is None else cc.size)
        self.lib.nbref_engine_fmaddn_inplace(self.h, a, self._marr(b), len(b), _ptr(cc), n)

    def fmaddn_corr(self, a, corr, b, c, csize=None):
        cc = None if c is None else np.ascontiguousarray(c, dtype=self.lib.dtype)
        n = csize if csize is not None else (0 if cc is None else cc.size)
        self.lib.nbref_engine_fmaddn_corr(self.h, a, corr, self._marr(b), len(b), _ptr(cc), n)

    def fmaddn(self, a, b, c, d, dsize=None):
        dd = None if d is None else np.ascontiguousarray(d, dtype=self.lib.dtype)
        n = dsize if dsize is not None else (0 if dd is None else dd.size)
        self.lib.nbref_engine_fmaddn(self.h, a, b, self._marr(c), len(c), _ptr(dd), n)

    def fmaxabs(self, a, garbage=2878767678687.0):
        out = np.array([garbage], dtype=self.lib.dtype)
        self.lib.nbref_engine_fmaxabs(self.h, a, _ptr(out))
        return out[0]

    def set_step(self, s):
        self.lib.nbref_engine_set_step(self.h, s)

    def get_step(self):
        return self.lib.nbref_engine_get_step(self.h)

    def get_time(self):
        return self.lib.nbref_engine_get_time(self.h)

    def compute_count(self):
        return self.lib.nbref_engine_compute_count(self.h)

    def fcompute_y(self):
        """f(get_y()) as a host array (fresh scratch buffer)."""
        f = self.create_buffer(self.problem_size() * np.dtype(self.lib.dtype).itemsize)
        self.fcompute(0.0, self.get_y(), f)
        out = self.read_buffer(f)
        self.free_buffer(f)
        return out

    def time_fcompute(self, reps=1):
        return self.lib.nbref_engine_time_fcompute(self.h, reps)


class Solver:
    def __init__(self, lib, **kw):
        self.lib = lib
        self.h = lib.nbref_solver_create(params(**kw))
        if not self.h:
            raise ValueError("solver factory returned NULL for %r" % (kw,))

    def close(self):
        """Must be called BEFORE the engine is closed (solver dtors free buffers via engine())."""
        if self.h:
            self.lib.nbref_solver_free(self.h)
            self.h = None

    def set_engine(self, e):
        self.lib.nbref_solver_set_engine(self.h, e.h)

    def set_time_step(self, mn, mx):
        self.lib.nbref_solver_set_time_step(self.h, mn, mx)

    def advise(self, dt):
        self.lib.nbref_solver_advise(self.h, dt)

    def run(self, data, max_time, dump_dt=0.0, check_dt=0.0):
        return self.lib.nbref_solver_run(self.h, data.h, max_time, dump_dt, check_dt)

    def butcher_check(self):
        out = np.zeros(3)
        if self.lib.nbref_solver_butcher_check(self.h, _ptr(out)) != 0:
            return None
        return out


class Heap:
    """nbody_space_heap(_stackless) build + export + instrumented walk."""

    def __init__(self, lib):
        self.lib = lib
        self.h = lib.nbref_heap_new()
        self.n = 0

    def close(self):
        if self.h:
            self.lib.nbref_heap_free(self.h)
            self.h = None

    def build(self, y, mass, ratio):
        y = np.ascontiguousarray(y, dtype=self.lib.dtype)
        mass = np.ascontiguousarray(mass, dtype=self.lib.dtype)
        self.n = mass.size
        return self.lib.nbref_heap_build(self.h, self.n, _ptr(y), _ptr(mass), ratio)

    def rebuild(self, y, ratio):
        y = np.ascontiguousarray(y, dtype=self.lib.dtype)
        self.lib.nbref_heap_rebuild(self.h, self.n, _ptr(y), ratio)

    def export(self):
        ts = 2 * self.n
        xyzr = np.zeros(4 * ts, dtype=self.lib.dtype)
        mass = np.zeros(ts, dtype=self.lib.dtype)
        body = np.zeros(ts, dtype=np.int64)
        got = self.lib.nbref_heap_export(self.h, _ptr(xyzr), _ptr(mass), _ptr(body))
        assert got == ts
        return xyzr.reshape(ts, 4), mass, body

    def walk_counts(self, first=0, last=None, stride=1):
        out = np.zeros(2, dtype=np.uint64)
        self.lib.nbref_heap_walk_counts(self.h, first, self.n if last is None else last, stride, _ptr(out))
        return int(out[0]), int(out[1])
