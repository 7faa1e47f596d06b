"""the get_data side of the engine at N = 4,194,304 (FP64: 192 MiB of positions + velocities per
frame). Three ways to land the state in nbody_data's AoS arrays, wall time per frame:
  read_buffer + host transpose   what nbody_engine_cuda::get_data does (full-buffer D2H into a temporary, then a host
                                 loop over N bodies, nbody_engine_cuda.cpp:141-175); the host loop is numpy here
  read_bodies, pageable          device transpose + two D2H copies per shard into ordinary memory
  read_bodies, registered        the same into arrays pinned once with nb200_host_register (what the adapter does)
Run on the GPU box:  python profiles/measure_get_data.py > profiles/r1_get_data.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbody_b200 import Engine  # noqa: E402


def timed(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    rng = np.random.RandomState(1)
    pos, vel = rng.standard_normal((n, 3)), rng.standard_normal((n, 3))
    mass = np.ones(n)
    out = {"bodies": n, "frame_bytes": int(pos.nbytes + vel.nbytes)}
    with Engine(precision="f64", devices="0") as e:
        assert e.init_bodies(pos, vel, mass)
        y = e.get_y()
        tmp = np.empty(6 * n)
        p0, v0 = np.empty((n, 3)), np.empty((n, 3))

        def old():
            e.read_into(tmp.ctypes.data, y)
            rows = tmp.reshape(6, n)
            p0[:] = rows[0:3].T
            v0[:] = rows[3:6].T
        out["read_buffer_plus_host_transpose_s"] = timed(old)
        p1, v1 = np.empty((n, 3)), np.empty((n, 3))
        out["read_bodies_pageable_s"] = timed(lambda: e.get_bodies(p1, v1))
        p2, v2 = np.empty((n, 3)), np.empty((n, 3))
        assert e.host_register(p2) == 0 and e.host_register(v2) == 0
        out["read_bodies_registered_s"] = timed(lambda: e.get_bodies(p2, v2))
        out["registered_GBps"] = out["frame_bytes"] / out["read_bodies_registered_s"] / 1e9
        out["identical"] = bool(np.array_equal(p0, pos) and np.array_equal(p1, pos) and np.array_equal(p2, pos)
                                and np.array_equal(v0, vel) and np.array_equal(v1, vel) and np.array_equal(v2, vel))
        e.host_unregister(p2)
        e.host_unregister(v2)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
