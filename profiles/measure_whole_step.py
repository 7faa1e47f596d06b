"""Whole solver steps in the reference's own process model (ONE process, --device=0,1,...: lanes), through the reference's
simulation flow with both overlay patches (oracle/_ref/nbody_sim_f64):

  BASELINE configs[2]  rkdp (7 fcompute + fused fmaddn stages + one fmaxabs per step), N = 1,048,576, 1 / 2 / 4 / 8 GPUs
  BASELINE configs[4]  rkfeagin14 (35 stages, fmaddn_corr with correction=1) and Bulirsch-Stoer, N = 262,144

plus the latency of the host-visible reduction (fmaxabs) and of a fused 7-term stage per lane count. The state ops'
share of a step is (step - fcompute_calls x fcompute) / step with the fcompute time measured in the same process model.

    python profiles/measure_whole_step.py [out.json] [--quick]
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
from sim_util import run_sim
from util import universe

out = [a for a in sys.argv[1:] if not a.startswith("--")]
out = out[0] if out else os.path.join(ROOT, "gpurun_out", "r2_whole_step.json")
quick = "--quick" in sys.argv
ndev = device_count()
counts = [g for g in (1, 2, 4, 8) if g <= ndev]
only = [a for a in sys.argv[1:] if a.startswith("--lanes=")]
if only:
    counts = [int(v) for v in only[0][8:].split(",")]
res = {"devices_present": ndev}


def devlist(g):
    return ",".join(str(i) for i in range(g))


# ---- state-op latencies per lane count (Python mirror of the same C ABI, lanes mode) ----
for n in (1 << 20, 1 << 18):
    y, m = universe(n)
    for g in counts:
        for transport in (("peer",) if g == 1 else ("peer", "nccl")):
            with Engine(devices=devlist(g)) as e:
                if transport == "nccl":
                    assert e.set_option("use_nccl", 1) == 0, e.last_error()
                assert e.init(y, m)
                bufs = e.create_buffers(e.get_y().size(), 8)
                for b in bufs:
                    e.copy_buffer(b, e.get_y())
                coeff = np.linspace(0.1, 0.7, 7)
                f = bufs[7]
                e.fcompute(0, e.get_y(), f)
                e.synchronize()
                reps = 20
                t0 = time.perf_counter()
                for _ in range(reps):
                    e.fmaxabs(bufs[0])
                t_max = (time.perf_counter() - t0) / reps
                e.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    e.fmaddn(bufs[7], bufs[0], bufs[:7], coeff)
                e.synchronize()
                t_fm = (time.perf_counter() - t0) / reps
                k = 1 if n > (1 << 18) else 3
                t0 = time.perf_counter()
                for _ in range(k):
                    e.fcompute(0, e.get_y(), f)
                e.synchronize()
                t_fc = (time.perf_counter() - t0) / k
                key = "ops_n%d_lanes%d_%s" % (n, g, transport)
                res[key] = {"fmaxabs_us": t_max * 1e6, "fmaddn7_us": t_fm * 1e6, "fcompute_ms": t_fc * 1e3}
                print(key, res[key], flush=True)

# ---- whole steps through the reference's solver classes ----
# max_recursion=0: every step is accepted after its embedded error estimate (fmaddn + fmaxabs) has been computed. With the
# factory default (8 levels of 8 sub-steps) one "step" of this galaxy costs hundreds of fcompute calls -- what is timed
# here is the unit all of them are made of: 7 stages + fused state ops + one host-visible max-norm.
runs = [("C3_rkdp_n1m", dict(solver="rkdp", stars_count=524288, max_recursion=0), 2 if quick else 3),
        ("C5_rkfeagin14_corr_n262144", dict(solver="rkfeagin14", stars_count=131072, correction=1, max_recursion=0), 2 if quick else 3),
        ("C5_bs_n262144", dict(solver="bs", stars_count=131072, max_level=4), 2 if quick else 3)]
for name, cfg, steps in runs:
    for g in counts:
        for transport in (("peer",) if g == 1 else ("peer", "nccl")):
            opts = dict(cfg, engine="b200", device=devlist(g), use_nccl=1 if transport == "nccl" else 0, max_steps=steps,
                        check_step=0, max_time=1e9, initial_type="G1")
            rows, summary, _ = run_sim(timeout=3000, **opts)
            n = summary["bodies"]
            fc = res["ops_n%d_lanes%d_%s" % (n, g, transport)]["fcompute_ms"]
            calls, calls1 = summary["fcompute_calls"], summary["fcompute_calls_first_step"]
            wall_ms, first_ms = summary["wall_s"] * 1e3, summary["first_step_s"] * 1e3
            key = "%s_lanes%d_%s" % (name, g, transport)
            # the first solver step creates the solver's buffers and the engine's scratch (one-time); rates are taken over
            # the steps after it. `steps` is the engine's step counter (Bulirsch-Stoer advances it per inner midpoint step).
            steady_ms, steady_calls = wall_ms - first_ms, calls - calls1
            res[key] = {"bodies": n, "steps": summary["steps"], "fcompute_calls": calls, "wall_ms": wall_ms, "first_step_ms": first_ms,
                        "fcompute_calls_after_first_step": steady_calls, "ms_per_fcompute_call_after_first_step": steady_ms / steady_calls,
                        "fcompute_ms_alone": fc, "state_ops_and_host_share": max(0.0, 1.0 - steady_calls * fc / steady_ms),
                        "pairs_per_s_whole_step": steady_calls * float(n) * n / (steady_ms * 1e-3)}
            print(key, res[key], flush=True)
json.dump(res, open(out, "w"), indent=1)
