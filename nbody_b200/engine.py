"""Python host mirror of the reference's ``nbody_engine`` API over the nb200 C ABI.

The reference's host language is C++ (the real adapter is
``nbody_b200/host/nbody_engine_b200.cpp``); this module offers the same
operator surface -- same method names, argument meaning and error behaviour as
``nbody/nbody_engine.h:16-96`` -- to Python callers (``bench.py``, the parity
tests). Everything goes through ``include/nb200.h`` via ctypes. There is no CPU
path: if the CUDA library is missing or no GPU is present, construction fails.

Error convention (``nbody/nbody_engine_cuda.cpp:205-217``): a call with a
foreign / NULL / mis-sized buffer logs one line and returns without touching
anything; it never raises. Only construction and ``init`` report failure.
"""
import ctypes as C
import logging
import os

import numpy as np

from . import build as _build

log = logging.getLogger("nb200")

TREE_LAYOUTS = {"heap": 1, "heap_stackless": 2}
UID_BYTES = 128

_LIBS = {}


class NativeLibraryMissing(RuntimeError):
    pass


def load_library(precision="f64"):
    """dlopen ``libnb200_<precision>.so`` and type every symbol of include/nb200.h."""
    if precision in _LIBS:
        return _LIBS[precision]
    path = _build.lib_path(precision)
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            "%s not found: build it with `python -m nbody_b200.build` (there is no CPU fallback)" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    real = C.c_double if precision == "f64" else C.c_float
    vp, sz, i32, ull = C.c_void_p, C.c_size_t, C.c_int, C.c_ulonglong
    P = C.POINTER

    def sig(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("nb200_real_size", i32)
    sig("nb200_device_count", i32, P(i32))
    sig("nb200_comm_unique_id", i32, vp)
    sig("nb200_create", i32, P(vp), P(i32), i32, i32, i32, vp)
    sig("nb200_destroy", i32, vp)
    sig("nb200_last_error", C.c_char_p, vp)
    sig("nb200_sync", i32, vp)
    sig("nb200_shards", i32, vp, P(i32), P(i32))
    sig("nb200_describe", i32, vp, C.c_char_p, sz)
    sig("nb200_set_bodies", i32, vp, sz, vp)
    sig("nb200_get_mass", i32, vp, vp)
    sig("nb200_alloc", i32, vp, sz, P(vp))
    sig("nb200_free", i32, vp, vp)
    sig("nb200_size", sz, vp)
    sig("nb200_write", i32, vp, vp, vp)
    sig("nb200_read", i32, vp, vp, vp)
    sig("nb200_read_local", i32, vp, vp, vp)
    sig("nb200_copy", i32, vp, vp, vp)
    sig("nb200_fill", i32, vp, vp, real)
    sig("nb200_lane_ptr", i32, vp, vp, i32, P(vp), P(sz))
    sig("nb200_fcompute_direct", i32, vp, vp, vp)
    sig("nb200_bh_configure", i32, vp, real, i32, sz)
    sig("nb200_fcompute_bh", i32, vp, vp, vp, sz)
    sig("nb200_bh_export_tree", i32, vp, i32, vp, vp, vp)
    sig("nb200_bh_walk_stats", i32, vp, i32, P(ull), P(ull))
    sig("nb200_bh_walk_profile", i32, vp, P(ull))
    sig("nb200_fmadd_inplace", i32, vp, vp, vp, real)
    sig("nb200_fmadd", i32, vp, vp, vp, vp, real)
    sig("nb200_fmaddn_inplace", i32, vp, vp, P(vp), vp, sz)
    sig("nb200_fmaddn", i32, vp, vp, vp, P(vp), vp, sz)
    sig("nb200_fmaddn_corr", i32, vp, vp, vp, P(vp), vp, sz)
    sig("nb200_fmaxabs", i32, vp, vp, P(real))
    sig("nb200_clamp", i32, vp, vp, real)
    sig("nb200_statistics", i32, vp, vp, i32, P(C.c_double))
    sig("nb200_write_bodies", i32, vp, vp, vp, vp)
    sig("nb200_read_bodies", i32, vp, vp, vp, vp)
    sig("nb200_host_register", i32, vp, vp, sz)
    sig("nb200_host_unregister", i32, vp, vp)
    sig("nb200_step_boundary", i32, vp)
    sig("nb200_step_graph_stats", i32, vp, P(ull))
    sig("nb200_launch_count", ull, vp)
    sig("nb200_last_fcompute_ms", i32, vp, P(C.c_float))
    sig("nb200_last_direct_path", i32, vp)
    sig("nb200_mark", i32, vp, i32)
    sig("nb200_elapsed_ms", i32, vp, i32, i32, P(C.c_float))
    sig("nb200_probe_fma_peak", i32, vp, C.c_double, P(C.c_double))
    sig("nb200_set_option", i32, vp, C.c_char_p, C.c_longlong)
    lib.real = real
    lib.dtype = np.dtype(np.float64 if precision == "f64" else np.float32)
    lib.precision = precision
    lib.path = path
    if lib.nb200_real_size() != lib.dtype.itemsize:
        raise NativeLibraryMissing("%s was built for another precision" % path)
    _LIBS[precision] = lib
    return lib


def device_count(precision="f64"):
    lib = load_library(precision)
    n = C.c_int(0)
    lib.nb200_device_count(C.byref(n))
    return n.value


def parse_devices(text, count):
    """``select_devices`` of the reference (nbody_engine_cuda.cpp:576-616):
    comma list of ids, every id in [0, count); anything else -> None."""
    parts = [p for p in str(text).split(",") if p != ""]
    if not parts:
        log.error("CUDA device list is empty")
        return None
    ids = []
    for p in parts:
        try:
            v = int(p.strip(), 10)
        except ValueError:
            log.error("Can't parse device ID %r", p)
            return None
        if v < 0 or v >= count:
            log.error("Invalid device ID %d must be in range [0 ... %d)", v, count)
            return None
        ids.append(v)
    return ids


def new_unique_id(precision="f64"):
    """128-byte NCCL id (rank 0 creates it; the launcher broadcasts it)."""
    lib = load_library(precision)
    buf = C.create_string_buffer(UID_BYTES)
    if lib.nb200_comm_unique_id(buf) != 0:
        raise RuntimeError("nb200_comm_unique_id failed (NCCL not loadable?)")
    return buf.raw


class Memory:
    """``nbody_engine::memory``: opaque device buffer of ``size()`` bytes."""

    __slots__ = ("handle", "_size", "engine")

    def __init__(self, engine, handle, size):
        self.engine = engine
        self.handle = handle
        self._size = size

    def size(self):
        return self._size


class Engine:
    """B200 engine with the method surface of ``nbody_engine`` (nbody_engine.h:37-88).

    ``kind`` selects the right-hand side the way the reference's factory aliases do
    (nbody/nbody_engines.cpp:21-83): ``"direct"`` ~ ``cuda``; ``"bh"`` ~ ``cuda_bh_tex``
    with ``distance_to_node_radius_ratio``, ``tree_build_rate``, ``tree_layout``.
    """

    def __init__(self, precision="f64", devices="0", rank=0, nranks=1, uid=None, kind="direct",
                 distance_to_node_radius_ratio=10.0, tree_build_rate=0, tree_layout="heap_stackless"):
        self.lib = load_library(precision)
        self.dtype = self.lib.dtype
        self.kind = kind
        if kind not in ("direct", "bh"):
            raise ValueError("kind must be 'direct' or 'bh'")
        if tree_layout not in TREE_LAYOUTS:
            raise ValueError("Invalid tree_layout. Allowed values are 'heap' or 'heap_stackless'")
        ids = devices if isinstance(devices, (list, tuple)) else parse_devices(devices, device_count(precision))
        if ids is None:
            raise ValueError("invalid device list %r" % (devices,))
        arr = (C.c_int * len(ids))(*ids)
        self.ctx = C.c_void_p()
        self._uid = C.create_string_buffer(uid, UID_BYTES) if uid is not None else None
        rc = self.lib.nb200_create(C.byref(self.ctx), arr, len(ids), rank, nranks, self._uid)
        if rc != 0:
            self.ctx = C.c_void_p()
            raise RuntimeError("nb200_create failed (%d): no usable CUDA device / NCCL; there is no CPU fallback" % rc)
        self.rank, self.nranks = rank, nranks
        self._n = 0
        self._y = None
        self._time = 0.0
        self._step = 0
        self._compute_count = 0
        if kind == "bh":
            self._check(self.lib.nb200_bh_configure(self.ctx, distance_to_node_radius_ratio,
                                                    TREE_LAYOUTS[tree_layout], tree_build_rate), "bh_configure")

    # ---- life cycle -------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.nb200_destroy(self.ctx)
            self.ctx = C.c_void_p()
            self._y = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        """Log-and-return, like the reference's qDebug() << "... is not smemory"; return;"""
        if rc != 0:
            log.warning("%s: %s", what, self.lib.nb200_last_error(self.ctx).decode(errors="replace"))
        return rc

    def _h(self, m, what, name):
        if not isinstance(m, Memory) or m.engine is not self or not m.handle:
            log.warning("%s: %s is not a buffer of this engine", what, name)
            return None
        return m.handle

    def type_name(self):
        return "nbody_engine_b200" if self.kind == "direct" else "nbody_engine_b200_bh"

    def last_error(self):
        return self.lib.nb200_last_error(self.ctx).decode(errors="replace")

    def print_info(self):
        buf = C.create_string_buffer(4096)
        self.lib.nb200_describe(self.ctx, buf, len(buf))
        return buf.value.decode()

    # ---- init / data ------------------------------------------------------
    def init(self, y, mass):
        """``init(nbody_data*)``: y = [rx|ry|rz|vx|vy|vz] (6N), mass (N). False on failure."""
        mass = np.ascontiguousarray(mass, dtype=self.dtype)
        y = np.ascontiguousarray(y, dtype=self.dtype)
        if y.size != 6 * mass.size or mass.size == 0:
            log.error("init: y must hold 6*N values")
            return False
        if self._check(self.lib.nb200_set_bodies(self.ctx, mass.size, mass.ctypes.data_as(C.c_void_p)), "init") != 0:
            return False
        self._n = mass.size
        self._y = self.create_buffer(y.nbytes)
        if self._y is None:
            return False
        self.write_buffer(self._y, y)
        return True

    def get_data(self):
        """Host copy of the current state vector (``get_data`` without the AoS transpose)."""
        return self.read_buffer(self._y)

    def init_bodies(self, pos, vel, mass):
        """``init(nbody_data*)`` from the AoS body arrays themselves: pos, vel are N x 3 (``nbody_data::get_vertites()``
        / ``get_velosites()``); the transpose to [rx|ry|rz|vx|vy|vz] happens on the device (``nb200_write_bodies``)."""
        mass = np.ascontiguousarray(mass, dtype=self.dtype)
        pos = np.ascontiguousarray(pos, dtype=self.dtype)
        vel = np.ascontiguousarray(vel, dtype=self.dtype)
        if mass.size == 0 or pos.shape != (mass.size, 3) or vel.shape != (mass.size, 3):
            log.error("init_bodies: pos and vel must be N x 3")
            return False
        if self._check(self.lib.nb200_set_bodies(self.ctx, mass.size, mass.ctypes.data_as(C.c_void_p)), "init") != 0:
            return False
        self._n = mass.size
        self._y = self.create_buffer(6 * mass.size * self.dtype.itemsize)
        if self._y is None:
            return False
        return self._check(self.lib.nb200_write_bodies(self.ctx, self._y.handle, pos.ctypes.data_as(C.c_void_p),
                                                       vel.ctypes.data_as(C.c_void_p)), "init_bodies") == 0

    def get_bodies(self, pos=None, vel=None, y=None):
        """``get_data(nbody_data*)``: the state vector back as AoS body arrays (N x 3 each), transposed on the device
        and copied shard by shard straight into ``pos`` / ``vel`` (pin them once with ``host_register``)."""
        y = self._y if y is None else y
        h = self._h(y, "get_bodies", "y")
        if h is None:
            return None
        pos = np.empty((self._n, 3), dtype=self.dtype) if pos is None else pos
        vel = np.empty((self._n, 3), dtype=self.dtype) if vel is None else vel
        for a in (pos, vel):
            if a.dtype != self.dtype or a.shape != (self._n, 3) or not a.flags.c_contiguous:
                log.warning("get_bodies: destination must be a contiguous N x 3 array of the engine's dtype")
                return None
        if self._check(self.lib.nb200_read_bodies(self.ctx, h, pos.ctypes.data_as(C.c_void_p),
                                                  vel.ctypes.data_as(C.c_void_p)), "get_bodies") != 0:
            return None
        return pos, vel

    def host_register(self, array):
        return self._check(self.lib.nb200_host_register(self.ctx, array.ctypes.data_as(C.c_void_p), array.nbytes),
                           "host_register")

    def host_unregister(self, array):
        return self.lib.nb200_host_unregister(self.ctx, array.ctypes.data_as(C.c_void_p))

    def problem_size(self):
        return 6 * self._n

    def get_y(self):
        return self._y

    def advise_time(self, dt):
        """End of a solver step; also the boundary the library's step graphs are cut at (``step_graph`` option)."""
        self._time += dt
        self._step += 1
        self._check(self.lib.nb200_step_boundary(self.ctx), "step_boundary")

    def step_graph_stats(self):
        out = (C.c_ulonglong * 5)()
        self.lib.nb200_step_graph_stats(self.ctx, out)
        return dict(graph_launches=int(out[0]), bailouts=int(out[1]),
                    state=("off", "record", "capture", "replay")[int(out[2])], launches_per_step=int(out[3]),
                    distinct_steps=int(out[4]))

    def get_time(self):
        return self._time

    def set_time(self, t):
        self._time = t

    def get_step(self):
        return self._step

    def set_step(self, s):
        self._step = s

    def advise_compute_count(self):
        self._compute_count += 1

    def get_compute_count(self):
        return self._compute_count

    # ---- buffers ----------------------------------------------------------
    def create_buffer(self, nbytes):
        h = C.c_void_p()
        if self._check(self.lib.nb200_alloc(self.ctx, nbytes, C.byref(h)), "create_buffer") != 0:
            return None
        return Memory(self, h, nbytes)

    def create_buffers(self, nbytes, count):
        out = []
        for _ in range(count):
            m = self.create_buffer(nbytes)
            if m is None:
                self.free_buffers(out)
                return []
            out.append(m)
        return out

    def free_buffer(self, m):
        if m is None:
            return
        h = self._h(m, "free_buffer", "m")
        if h is not None:
            self.lib.nb200_free(self.ctx, h)
            m.handle = None

    def free_buffers(self, ms):
        for m in ms:
            self.free_buffer(m)
        del ms[:]

    def write_buffer(self, dst, src):
        h = self._h(dst, "write_buffer", "dst")
        if h is None:
            return
        if src is None:
            log.warning("write_buffer: NULL source")
            return
        src = np.ascontiguousarray(src)
        if src.nbytes < dst.size():
            log.warning("write_buffer: host array smaller than the buffer")
            return
        self._check(self.lib.nb200_write(self.ctx, h, src.ctypes.data_as(C.c_void_p)), "write_buffer")

    def read_buffer(self, src, dtype=None):
        h = self._h(src, "read_buffer", "src")
        if h is None:
            return None
        dtype = np.dtype(dtype or self.dtype)
        out = np.empty(src.size() // dtype.itemsize, dtype=dtype)
        if out.size:
            self._check(self.lib.nb200_read(self.ctx, out.ctypes.data_as(C.c_void_p), h), "read_buffer")
        return out

    def read_into(self, host, src):
        """read_buffer into caller-owned (e.g. pinned) memory."""
        h = self._h(src, "read_buffer", "src")
        if h is not None:
            self._check(self.lib.nb200_read(self.ctx, C.c_void_p(host), h), "read_buffer")

    def read_local_into(self, host, src):
        """Only this process's shard columns of `src` into the full-layout host array (no gather across ranks)."""
        h = self._h(src, "read_buffer", "src")
        if h is not None:
            self._check(self.lib.nb200_read_local(self.ctx, C.c_void_p(host), h), "read_buffer")

    def write_from(self, dst, host):
        h = self._h(dst, "write_buffer", "dst")
        if h is not None:
            self._check(self.lib.nb200_write(self.ctx, h, C.c_void_p(host)), "write_buffer")

    def copy_buffer(self, a, b):
        ha, hb = self._h(a, "copy_buffer", "a"), self._h(b, "copy_buffer", "b")
        if ha is None or hb is None:
            return
        self._check(self.lib.nb200_copy(self.ctx, ha, hb), "copy_buffer")

    def fill_buffer(self, a, value):
        ha = self._h(a, "fill_buffer", "a")
        if ha is None:
            return
        self._check(self.lib.nb200_fill(self.ctx, ha, value), "fill_buffer")

    def lane_ptr(self, m, lane=0):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.nb200_lane_ptr(self.ctx, m.handle, lane, C.byref(p), C.byref(n)), "lane_ptr")
        return p.value, n.value

    # ---- f(t, y) ------------------------------------------------------------
    def fcompute(self, t, y, f):
        hy, hf = self._h(y, "fcompute", "y"), self._h(f, "fcompute", "f")
        if hy is None or hf is None:
            return
        self.advise_compute_count()
        if self.kind == "direct":
            self._check(self.lib.nb200_fcompute_direct(self.ctx, hy, hf), "fcompute")
        else:
            self._check(self.lib.nb200_fcompute_bh(self.ctx, hy, hf, self._step), "fcompute")

    def clamp(self, y, b):
        hy = self._h(y, "clamp", "y")
        if hy is None:
            return
        self._check(self.lib.nb200_clamp(self.ctx, hy, b), "clamp")

    # ---- state-vector ops -------------------------------------------------
    def fmadd_inplace(self, a, b, c):
        ha, hb = self._h(a, "fmadd_inplace", "a"), self._h(b, "fmadd_inplace", "b")
        if ha is None or hb is None:
            return
        self._check(self.lib.nb200_fmadd_inplace(self.ctx, ha, hb, c), "fmadd_inplace")

    def fmadd(self, a, b, c, d):
        ha, hb, hc = self._h(a, "fmadd", "a"), self._h(b, "fmadd", "b"), self._h(c, "fmadd", "c")
        if ha is None or hb is None or hc is None:
            return
        self._check(self.lib.nb200_fmadd(self.ctx, ha, hb, hc, d), "fmadd")

    def _terms(self, what, bufs, coeff, csize):
        """Coefficient array + handle array for the fused calls; None -> log and return."""
        if coeff is None:
            log.warning("%s: coefficient array is NULL", what)
            return None
        coeff = np.ascontiguousarray(coeff, dtype=self.dtype)
        n = coeff.size if csize is None else csize
        if n > len(bufs):
            log.warning("%s: csize > b.size()", what)
            return None
        if n > coeff.size:
            log.warning("%s: csize > len(c)", what)
            return None
        handles = (C.c_void_p * max(1, n))()
        for k in range(n):
            m = bufs[k]
            handles[k] = m.handle if isinstance(m, Memory) and m.engine is self else None
        return coeff, handles, n

    def fmaddn_inplace(self, a, b, c, csize=None):
        ha = self._h(a, "fmaddn_inplace", "a")
        t = self._terms("fmaddn_inplace", b, c, csize)
        if ha is None or t is None:
            return
        self._check(self.lib.nb200_fmaddn_inplace(self.ctx, ha, t[1], t[0].ctypes.data_as(C.c_void_p), t[2]), "fmaddn_inplace")

    def fmaddn(self, a, b, c, d, dsize=None):
        ha = self._h(a, "fmaddn", "a")
        hb = None
        if b is not None:
            hb = self._h(b, "fmaddn", "b")
            if hb is None:
                return
        t = self._terms("fmaddn", c, d, dsize)
        if ha is None or t is None:
            return
        self._check(self.lib.nb200_fmaddn(self.ctx, ha, hb, t[1], t[0].ctypes.data_as(C.c_void_p), t[2]), "fmaddn")

    def fmaddn_corr(self, a, corr, b, c, csize=None):
        ha, hc = self._h(a, "fmaddn_corr", "a"), self._h(corr, "fmaddn_corr", "corr")
        t = self._terms("fmaddn_corr", b, c, csize)
        if ha is None or hc is None or t is None:
            return
        self._check(self.lib.nb200_fmaddn_corr(self.ctx, ha, hc, t[1], t[0].ctypes.data_as(C.c_void_p), t[2]), "fmaddn_corr")

    def fmaxabs(self, a, default=None):
        """``fmaxabs(a, result)``: returns the result; ``default`` is what the caller's variable
        would keep when the call is rejected (the reference leaves ``result`` untouched)."""
        ha = self._h(a, "fmaxabs", "a")
        if ha is None:
            return default
        out = self.lib.real(0)
        if self._check(self.lib.nb200_fmaxabs(self.ctx, ha, C.byref(out)), "fmaxabs") != 0:
            return default
        return self.dtype.type(out.value)

    # ---- conservation report (device-side print_statistics sums) ---------------
    def statistics(self, y=None, with_energy=True):
        """dict(P[3], L[3], Ekin, Epot, C[3]) of the state vector `y` (default: get_y()); see nb200_statistics."""
        hy = self._h(y if y is not None else self._y, "statistics", "y")
        if hy is None:
            return None
        out = (C.c_double * 11)()
        if self._check(self.lib.nb200_statistics(self.ctx, hy, 1 if with_energy else 0, out), "statistics") != 0:
            return None
        v = np.array(out[:], dtype=np.float64)
        return dict(P=v[0:3], L=v[3:6], Ekin=v[6], Epot=v[7], C=v[8:11])

    # ---- instrumentation ------------------------------------------------------
    def synchronize(self):
        self._check(self.lib.nb200_sync(self.ctx), "sync")

    def launch_count(self):
        return int(self.lib.nb200_launch_count(self.ctx))

    def last_fcompute_ms(self):
        out = (C.c_float * 4)()
        self.lib.nb200_last_fcompute_ms(self.ctx, out)
        return dict(pack_gather=out[0], tree=out[1], force=out[2], reduce=out[3])

    def last_direct_path(self):
        """0 = ordered-pair kernel, otherwise the tile edge of the symmetric-tile kernel."""
        return int(self.lib.nb200_last_direct_path(self.ctx))

    def mark(self, slot):
        """Record CUDA event `slot` on the engine's stream (device-side stopwatch)."""
        self._check(self.lib.nb200_mark(self.ctx, slot), "mark")

    def elapsed_ms(self, slot_a, slot_b):
        out = C.c_float(0)
        self._check(self.lib.nb200_elapsed_ms(self.ctx, slot_a, slot_b, C.byref(out)), "elapsed_ms")
        return out.value

    def probe_fma_peak(self, ms=200.0):
        out = C.c_double(0)
        self._check(self.lib.nb200_probe_fma_peak(self.ctx, ms, C.byref(out)), "probe_fma_peak")
        return out.value

    def set_option(self, name, value):
        return self._check(self.lib.nb200_set_option(self.ctx, name.encode(), int(value)), "set_option")

    def shards(self):
        a, b = C.c_int(0), C.c_int(0)
        self.lib.nb200_shards(self.ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def bh_export_tree(self, lane=0):
        ts = 2 * self._n
        xyzr = np.zeros((ts, 4), dtype=self.dtype)
        mass = np.zeros(ts, dtype=self.dtype)
        body = np.zeros(ts, dtype=np.int32)
        rc = self._check(self.lib.nb200_bh_export_tree(self.ctx, lane, xyzr.ctypes.data_as(C.c_void_p),
                                                       mass.ctypes.data_as(C.c_void_p), body.ctypes.data_as(C.c_void_p)),
                         "bh_export_tree")
        return (xyzr, mass, body) if rc == 0 else None

    def bh_walk_profile(self):
        """Counters of the last counted grouped walk (see nb200_bh_walk_profile)."""
        out = (C.c_ulonglong * 16)()
        self.lib.nb200_bh_walk_profile(self.ctx, out)
        keys = ("rounds", "items", "entries", "unsure_lane_items", "clamp_rounds", "max_stack", "busiest_target_entries",
                "busiest_target_entries_whole_walk", "entries_32", "entries_24_31", "entries_16_23", "entries_8_15", "entries_1_7")
        return dict(zip(keys, [int(v) for v in out[:13]]))

    def bh_walk_stats(self, enable=True):
        v, k = C.c_ulonglong(0), C.c_ulonglong(0)
        self.lib.nb200_bh_walk_stats(self.ctx, 1 if enable else 0, C.byref(v), C.byref(k))
        return v.value, k.value
