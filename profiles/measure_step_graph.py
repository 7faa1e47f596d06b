"""whole solver steps of the reference's own solver classes (unmodified, oracle/_ref) on
nbody_engine_b200, issued call by call (step_graph=0) and replayed as CUDA graphs (step_graph=1). Wall time per step
over `steps` steps after warm-up, synchronised at both ends; the graph path must give the same state bit for bit.
Run on the GPU box:  python profiles/measure_step_graph.py > profiles/r1_step_graph.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")
from nbody_b200 import build  # noqa: E402
from oracle import refharness as R  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 16384, 65536]
    lib = R.load("f64")
    ad = C.CDLL(build.adapter_path("f64"))
    ad.nbody_engine_b200_create.restype = C.c_void_p
    ad.nbody_engine_b200_create.argtypes = [C.c_char_p]
    ad.nbody_engine_b200_launch_count.restype = C.c_ulonglong
    ad.nbody_engine_b200_launch_count.argtypes = [C.c_void_p]
    ad.nbody_engine_b200_synchronize.argtypes = [C.c_void_p]
    ad.nbody_engine_b200_step_graph_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    cases = [("euler", dict(solver="euler")),
             ("rk4", dict(solver="rk4")),
             ("rkdp", dict(solver="rkdp", max_recursion=1, error_threshold=1e10)),
             ("rkfeagin14", dict(solver="rkfeagin14", max_recursion=1, error_threshold=1e10)),
             ("bs (max_level 4)", dict(solver="bs", max_level=4, error_threshold=1e-30))]
    out = {"what": "wall seconds per solver step, reference solver classes on nbody_engine_b200, FP64", "rows": []}
    for n in sizes:
        for name, p in cases:
            row = {"bodies": n, "solver": name}
            states = {}
            for graph in (0, 1):
                d = R.Data(lib).make_universe(n // 2)
                h = ad.nbody_engine_b200_create(("engine=b200;device=0;step_graph=%d" % graph).encode())
                e = R.Engine(lib, handle=h)
                assert e.init(d)
                s = R.Solver(lib, **p)
                s.set_time_step(1e-9, 1e-3)
                s.set_engine(e)
                for _ in range(30 if n <= 16384 else 6):     # allocation, recording, capture, first replays (Bulirsch-Stoer and
                    # Adams need a few outer steps until every sub-step / buffer rotation has its graph)
                    s.advise(1e-3)
                ad.nbody_engine_b200_synchronize(h)
                steps = 50 if n <= 16384 else 10
                cc0, l0 = e.compute_count(), ad.nbody_engine_b200_launch_count(h)
                t0 = time.perf_counter()
                for _ in range(steps):
                    s.advise(1e-3)
                ad.nbody_engine_b200_synchronize(h)
                dt = (time.perf_counter() - t0) / steps
                st = (C.c_ulonglong * 5)()
                ad.nbody_engine_b200_step_graph_stats(h, st)
                key = "graph" if graph else "eager"
                row[key + "_s_per_step"] = dt
                row["fcompute_per_step"] = (e.compute_count() - cc0) // steps
                row["launches_per_step"] = (ad.nbody_engine_b200_launch_count(h) - l0) // steps
                if graph:
                    row["graphs_launched"] = int(st[0])
                    row["replays_abandoned"] = int(st[1])
                    row["state"] = ("off", "record", "capture", "replay")[int(st[2])]
                    row["distinct_steps"] = int(st[4])
                e.get_data(d)
                states[graph] = d.export()[0].copy()
                s.close()
                e.close()
                d.close()
            row["speedup"] = row["eager_s_per_step"] / row["graph_s_per_step"]
            row["bit_identical"] = bool(np.array_equal(states[0], states[1]))
            out["rows"].append(row)
            print(json.dumps(row), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
