"""Lanes mode (ONE process, several devices -- what `nbody-simulation --device=0,1,...` runs) at 1 / 2 / 4 / 8 GPUs:
direct fcompute at N = 1M and Barnes-Hut fcompute at N = 4M, shards exchanged by peer copies / peer loads (default)
and by in-process NCCL (use_nccl=1). Results are compared with the one-lane engine's.

    python profiles/measure_lanes.py [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from nbody_b200 import Engine, device_count
from util import universe
from conftest import rel_err_per_body

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_lanes.json")
res = {}
ndev = device_count()
lists = [",".join(str(i) for i in range(g)) for g in (1, 2, 4, 8) if g <= ndev]
for kind, n, steps in (("direct", 1 << 20, 3), ("bh", 1 << 22, 3)):
    y, m = universe(n)
    ref = None
    for devices in lists:
        for transport in ("peer", "nccl"):
            g = devices.count(",") + 1
            if g == 1 and transport == "nccl":
                continue
            with Engine(devices=devices, kind=kind, distance_to_node_radius_ratio=10.0) as e:
                if transport == "nccl":
                    assert e.set_option("use_nccl", 1) == 0, e.last_error()
                assert e.init(y, m), e.last_error()
                f = e.create_buffer(e.get_y().size())
                for _ in range(2):
                    e.fcompute(0, e.get_y(), f)
                e.synchronize()
                e.mark(0)
                t0 = time.perf_counter()
                for _ in range(steps):
                    e.fcompute(0, e.get_y(), f)
                e.mark(1)
                e.synchronize()
                wall = (time.perf_counter() - t0) / steps
                got = e.read_buffer(f)
                if ref is None:
                    ref = got
                err = rel_err_per_body(got, ref, n)
                key = "%s_n%d_lanes%d_%s" % (kind, n, g, transport)
                res[key] = {"devices": devices, "transport": transport, "ms_wall": wall * 1e3, "ms_lane0_device": e.elapsed_ms(0, 1) / steps,
                            "vs_one_lane_rel_err": err, "phases_ms": e.last_fcompute_ms()}
                if kind == "direct":
                    res[key]["pairs_per_s"] = float(n) * n / wall
                    res[key]["path"] = e.last_direct_path()
                print(key, res[key], flush=True)
                assert err <= (1e-13 if kind == "direct" else 0.0), (key, err)
json.dump(res, open(out, "w"), indent=1)
