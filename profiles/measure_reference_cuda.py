"""The reference's own CUDA kernels recompiled for sm_100a (oracle/_ref/libnbref_cuda_f64.so) timed on the same B200
next to nb200, on the same inputs, with results compared.
    python profiles/measure_reference_cuda.py > profiles/r1_reference_cuda.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NBREF_QUIET", "1")
from nbody_b200 import Engine  # noqa: E402
from oracle import refcuda  # noqa: E402
from util import universe  # noqa: E402


def rel(a, b, n):
    a, b = a.reshape(6, n)[3:], b.reshape(6, n)[3:]
    return float((np.sqrt(((a - b) ** 2).sum(0)) / np.sqrt((b ** 2).sum(0))).max())


def main():
    out = {"direct": [], "bh": []}
    n = 262144
    y, m = universe(n)
    with Engine(devices=[0]) as e:
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0, e.get_y(), f)
        e.fcompute(0, e.get_y(), f)
        mine_ms = e.last_fcompute_ms()
        mine = e.read_buffer(f)
    mine_total = sum(mine_ms.values())
    for bs in (64, 128, 256, 512, 1024):
        f_ref, ms = refcuda.direct(y, m, block_size=bs, reps=2)
        out["direct"].append({"bodies": n, "kernel": "reference kfcompute+kfcompute_xyz (sm_100a)", "block_size": bs, "ms": ms,
                              "pairs_per_s": float(n) * n / (ms * 1e-3), "nb200_ms": mine_total,
                              "nb200_pairs_per_s": float(n) * n / (mine_total * 1e-3), "rel_err_vs_nb200": rel(f_ref, mine, n)})
    for n, ratio in ((1 << 20, 10.0), (1 << 22, 10.0)):
        y, m = universe(n)
        with Engine(devices=[0], kind="bh", distance_to_node_radius_ratio=ratio) as e:
            assert e.init(y, m)
            f = e.create_buffer(e.get_y().size())
            e.fcompute(0, e.get_y(), f)
            e.fcompute(0, e.get_y(), f)
            ph = e.last_fcompute_ms()
            mine = e.read_buffer(f)
            xyzr, nm, body = e.bh_export_tree()
        for bs in (64, 256, 1024):
            f_ref, ms = refcuda.bh_stackless(y, xyzr, nm, body, block_size=bs, reps=1)
            out["bh"].append({"bodies": n, "ratio": ratio,
                              "kernel": "reference kfcompute_heap_bh_stackless (texture tree, sm_100a), walk only",
                              "block_size": bs, "ms": ms, "nb200_walk_ms": ph["force"], "nb200_build_ms": ph["tree"],
                              "rel_err_vs_nb200": rel(f_ref, mine, n)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
