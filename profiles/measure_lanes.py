import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from nbody_b200 import Engine, device_count
from util import universe
res = {}
for n in (262144, 1 << 20):
    y, m = universe(n)
    for devices in ["0", "0,0"] + (["0,1"] if device_count() >= 2 else []):
        with Engine(devices=devices) as e:
            assert e.init(y, m)
            f = e.create_buffer(e.get_y().size())
            for _ in range(2): e.fcompute(0, e.get_y(), f)
            e.synchronize()
            import time
            t0 = time.perf_counter()
            k = 3
            for _ in range(k): e.fcompute(0, e.get_y(), f)
            e.synchronize()
            dt = (time.perf_counter() - t0) / k
            res[f"n{n}_dev{devices}"] = {"ms": dt * 1e3, "pairs_per_s": n * (n - 1) / dt, "path": e.last_direct_path()}
            print(n, devices, res[f"n{n}_dev{devices}"], flush=True)
json.dump(res, open('/root/repo/gpurun_out/lanes_time.json', 'w'), indent=1)
