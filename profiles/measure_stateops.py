"""HBM roofline of the fused state-vector kernels (run on the GPU box):
    python profiles/measure_stateops.py > profiles/r1_stateops.json
bytes = (#distinct buffers read + written) x 6N x sizeof(T)   (SURVEY.md 8d), device time by CUDA events on the
engine's stream, inputs > L2 or L2 flushed between repetitions."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbody_b200 import Engine  # noqa: E402


def main():
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    out = {"peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json" if os.path.exists(peaks) else "fallback", "rows": []}
    for n in (262144, 1 << 20, 1 << 22):
        e = Engine(precision="f64", devices=[0])
        rng = np.random.RandomState(0)
        y = rng.rand(6 * n)
        assert e.init(y, np.ones(n))
        sz = e.get_y().size()
        a, b = e.create_buffer(sz), e.create_buffer(sz)
        ks = e.create_buffers(sz, 35)
        corr = e.create_buffer(sz)
        flush = e.create_buffer(512 << 20)
        e.write_buffer(a, y)
        e.write_buffer(b, y)
        for k in ks:
            e.copy_buffer(k, b)
        e.fill_buffer(corr, 0)
        c35 = np.linspace(0.1, 1.0, 35)

        def timed(fn, streams, reps=5):
            best = 1e30
            for _ in range(reps):
                e.fill_buffer(flush, 0)
                e.mark(0)
                fn()
                e.mark(1)
                best = min(best, e.elapsed_ms(0, 1))
            gb = streams * sz / 1e9
            return {"ms": best, "bytes": streams * sz, "gbs": gb / (best * 1e-3), "frac": gb / (best * 1e-3) / hbm}

        rows = {
            "fmadd_inplace (2R+1W)": timed(lambda: e.fmadd_inplace(a, b, 0.5), 3),
            "fmadd (2R+1W)": timed(lambda: e.fmadd(a, b, ks[0], 0.5), 3),
            "fmaddn k=7 (8R+1W)": timed(lambda: e.fmaddn(a, b, ks[:7], c35[:7]), 9),
            "fmaddn k=35 (36R+1W)": timed(lambda: e.fmaddn(a, b, ks, c35), 37),
            "fmaddn_inplace k=4 (5R+1W)": timed(lambda: e.fmaddn_inplace(a, ks[:4], c35[:4]), 6),
            "fmaddn_corr k=7 (9R+2W)": timed(lambda: e.fmaddn_corr(a, corr, ks[:7], c35[:7]), 11),
            "fmaxabs (1R)": timed(lambda: e.fmaxabs(a), 1),
            "copy_buffer (1R+1W)": timed(lambda: e.copy_buffer(a, b), 2),
            "fill_buffer (1W)": timed(lambda: e.fill_buffer(a, 1.0), 1),
        }
        out["rows"].append({"bodies": n, "state_bytes": sz, "ops": rows})
        e.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
