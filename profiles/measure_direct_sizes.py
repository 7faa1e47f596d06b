"""Direct fcompute across N (BASELINE configs C1, C2, C5 and the sizes between): pair interactions/s of the ordered-pair
kernel (direct_symmetric=0) and of the symmetric tiles (direct_symmetric=1, automatic tile edge), FP64 and FP32, device
time from CUDA events on the engine's stream (nb200_mark), best of `reps` after warm-up.
Run on the GPU box:  python profiles/measure_direct_sizes.py > profiles/r1_direct_sizes.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NBREF_QUIET", "1")
from nbody_b200 import Engine  # noqa: E402
from util import numpy_universe  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144]
    out = {"what": "pair interactions/s (N^2 per fcompute), one B200, device time", "rows": []}
    for precision in ("f64", "f32"):
        dtype = np.float64 if precision == "f64" else np.float32
        for n in sizes:
            y, m = numpy_universe(n)
            row = {"precision": precision, "bodies": n}
            for label, sym in (("ordered_pairs", 0), ("symmetric_tiles", 1), ("automatic", -1)):
                with Engine(precision=precision) as e:
                    e.set_option("direct_symmetric", sym)
                    assert e.init(y.astype(dtype), m.astype(dtype))
                    f = e.create_buffer(e.get_y().size())
                    for _ in range(3):
                        e.fcompute(0.0, e.get_y(), f)
                    best = 1e30
                    reps = 10 if n <= 65536 else 3
                    for _ in range(reps):
                        e.mark(0)
                        e.fcompute(0.0, e.get_y(), f)
                        e.mark(1)
                        best = min(best, e.elapsed_ms(0, 1))
                    row[label] = float(n) * n / (best * 1e-3)
                    row[label + "_ms"] = best
                    if sym != 0:
                        row[label + "_tile_edge"] = e.last_direct_path()
            out["rows"].append(row)
            print(json.dumps(row), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
