"""BASELINE config C5: one step of the reference's many-stage solvers (rkfeagin14: 35 stages; Bulirsch-Stoer) at
N = 262,144 driven UNCHANGED on nbody_engine_b200 (C++ adapter over the C ABI); wall time per step, fcompute count
and kernel launches. Run on the GPU box:  python profiles/measure_solver_step.py > profiles/r1_solver_step.json"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")
from nbody_b200 import build  # noqa: E402
from oracle import refharness as R  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    lib = R.load("f64")
    ad = C.CDLL(build.adapter_path("f64"))
    ad.nbody_engine_b200_create.restype = C.c_void_p
    ad.nbody_engine_b200_create.argtypes = [C.c_char_p]
    ad.nbody_engine_b200_launch_count.restype = C.c_ulonglong
    ad.nbody_engine_b200_launch_count.argtypes = [C.c_void_p]
    ad.nbody_engine_b200_synchronize.argtypes = [C.c_void_p]
    out = {"bodies": n, "rows": []}
    cases = [("rk4", dict(solver="rk4")),
             ("rkdp", dict(solver="rkdp", max_recursion=1, error_threshold=1e10)),
             ("rkfeagin14", dict(solver="rkfeagin14", max_recursion=1, error_threshold=1e10)),
             ("rkfeagin14-corr", dict(solver="rkfeagin14", max_recursion=1, error_threshold=1e10, correction="true")),
             ("bs (max_level 4)", dict(solver="bs", max_level=4, error_threshold=1e-30))]
    for name, p in cases:
        d = R.Data(lib).make_universe(n // 2)
        h = ad.nbody_engine_b200_create(b"engine=b200;device=0")
        e = R.Engine(lib, handle=h)
        assert e.init(d)
        s = R.Solver(lib, **p)
        s.set_time_step(1e-9, 1e-3)
        s.set_engine(e)
        s.advise(1e-3)                                   # warm-up step: allocates the k buffers
        ad.nbody_engine_b200_synchronize(h)
        cc0, l0 = e.compute_count(), ad.nbody_engine_b200_launch_count(h)
        t0 = time.perf_counter()
        steps = 2
        for _ in range(steps):
            s.advise(1e-3)
        ad.nbody_engine_b200_synchronize(h)
        dt = (time.perf_counter() - t0) / steps
        cc, launches = (e.compute_count() - cc0) // steps, (ad.nbody_engine_b200_launch_count(h) - l0) // steps
        # split of the step: the same number of fcompute calls alone, on the same state, timed the same way
        fbuf = e.create_buffer(e.size(e.get_y()))
        e.fcompute(0, e.get_y(), fbuf)
        ad.nbody_engine_b200_synchronize(h)
        t1 = time.perf_counter()
        for _ in range(cc):
            e.fcompute(0, e.get_y(), fbuf)
        ad.nbody_engine_b200_synchronize(h)
        t_fc = time.perf_counter() - t1
        e.free_buffer(fbuf)
        out["rows"].append({"solver": name, "s_per_step": dt, "fcompute_per_step": cc, "launches_per_step": launches,
                            "fcompute_s_per_step": t_fc, "state_ops_s_per_step": max(dt - t_fc, 0.0),
                            "state_ops_share": max(dt - t_fc, 0.0) / dt, "pairs_per_s": cc * float(n) * n / dt})
        s.close()
        e.close()
        d.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
