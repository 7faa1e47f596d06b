"""Every solver class of the reference (compiled unmodified, oracle/_ref) on nbody_engine_b200 under the stand-in CUDA
runtime (see drive.py): after a learning phase each solver step must be replayed as CUDA graphs -- no capture, no
instantiation and no eagerly issued kernel except the fmaxabs reduction of the error-controlled solvers.
Kernels do not run here, so nothing is said about numbers (fmaxabs reads 0, no solver ever subdivides); the GPU suite
(tests/test_stepgraph_gpu.py) holds the golden end states with and without graphs. Test infrastructure."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NBREF_QUIET", "1")
from nbody_b200 import build  # noqa: E402
from oracle import refharness as R  # noqa: E402
from test_solvers_gpu import CASES  # noqa: E402

MOCK = C.CDLL(sys.argv[1])
EAGER, CAPTURED, GRAPHS, REPLAYED, INSTANTIATED, COPIES, CAPTURES, ALLOCS = range(8)


def counters():
    out = (C.c_ulonglong * 8)()
    MOCK.mock_counters(out)
    return np.array(list(out), dtype=np.int64)


def main():
    lib = R.load("f64")
    ad = C.CDLL(build.adapter_path("f64"))
    ad.nbody_engine_b200_create.restype = C.c_void_p
    ad.nbody_engine_b200_create.argtypes = [C.c_char_p]
    ad.nbody_engine_b200_step_graph_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    # solvers whose step contains the fmaxabs of an embedded error estimate or of the extrapolation error
    reads_norm = {"bulirsch-stoer", "rkck", "rkdp", "rkdp-corr", "rkdverk", "rkf", "rkfeagin10", "rkfeagin10-corr", "rkfeagin12",
                  "rkfeagin14", "rklc"}
    for name, params in CASES:
        d = R.Data(lib).load(os.path.join(ROOT, "tests", "golden", "initial_state.txt"))
        h = ad.nbody_engine_b200_create(b"engine=b200;device=0")          # step graphs are the adapter's default
        e = R.Engine(lib, handle=h)
        assert e.init(d)
        s = R.Solver(lib, **params)
        s.set_time_step(1e-3, 3e-2)
        s.set_engine(e)
        for _ in range(40):                                               # learning: allocation, recording, captures
            s.advise(3e-2)
        before = counters()
        steps = 10
        for _ in range(steps):
            s.advise(3e-2)
        c = counters() - before
        st = (C.c_ulonglong * 5)()
        assert ad.nbody_engine_b200_step_graph_stats(h, st) == 0
        line = "%-16s graphs %4d, kernels replayed %5d, eager %3d, distinct steps %2d, replays abandoned %d" % (
            name, c[GRAPHS], c[REPLAYED], c[EAGER], st[4], st[1])
        assert c[CAPTURES] == 0 and c[INSTANTIATED] == 0 and c[CAPTURED] == 0, line
        assert int(st[2]) == 3, line                                       # the next step will be replayed too
        assert c[GRAPHS] >= steps and c[REPLAYED] >= 2 * steps, line
        assert c[EAGER] == (steps if name in reads_norm else 0), line      # only the reduction runs when called
        assert (st[4] >= 5) == (name in ("adams5", "adams5-corr", "bulirsch-stoer")), line   # periodic patterns
        print("ok " + line)
        s.close()
        e.close()
        d.close()
    print("all reference solvers replayed")


if __name__ == "__main__":
    main()
