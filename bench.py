#!/usr/bin/env python
"""nb200 benchmark: BASELINE.json's headline metric on its own configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl nb200|reference] [--workload direct|bh]

Workload (default, `direct`): BASELINE.json configs[2] -- direct all-pairs fcompute, N = 1,048,576 bodies
(G1 synthetic galaxy pair), FP64, target bodies sharded over the N GPUs (one process per GPU under torchrun,
packed source bodies all-gathered with NCCL inside every fcompute). A "step" is one fcompute pass over the
whole system: N^2 pair interactions (self pairs included, as the reference's block/cuda kernels count them).
`value` = N^2 * K / (device time of K steps, max over ranks), inputs resident in HBM.
`e2e`   = the same metric through the public engine API with HOST buffers: every step writes y from pinned
          host memory (write_buffer), runs fcompute and reads f back (read_buffer).
`roofline` = the all-pairs kernel against the FP64 FMA pipe: SURVEY.md 8(d)'s 18 FP64-pipe instruction slots
          per pair (36 flop) x pairs per launch / that kernel's CUDA-event time; `peak` is a DFMA probe kernel
          timed in the same process (MEASURED_PEAKS.json has no FP64 entry; said in `peak_source`).
`cpu_baseline` = the reference's own nbody_engine_openmp (oracle/_ref, else the C port) on the host cores,
          on a bounded sample (N = 32,768 of the same galaxy model), rank 0 at N = 1 only.

--impl reference times the reference's CPU engine alone (same metric/unit) on the bounded sample.
--workload bh measures the second headline number (Barnes-Hut fcompute ms/step at N = 4,194,304).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")

N_DIRECT = 1 << 20
N_BH = 1 << 22
N_CPU_SAMPLE = 131072        # cpu_baseline sample: 1.7e10 pairs, ~5-20 s of CPU work per fcompute on 16-64 cores
N_CPU_STEP = 65536           # --impl reference: bodies per timed step (4.3e9 pairs, ~1 s on 16 cores)
# FMA-pipe instruction slots per pair -- FP64: 18 (the kernel issues 17); FP32: 13 (+1 MUFU on XU)
SLOTS_PER_PAIR = {"f64": 18, "f32": 13}
ISSUED_PER_PAIR = {"f64": 17, "f32": 13}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nb200", choices=["nb200", "reference"])
    ap.add_argument("--workload", default="direct", choices=["direct", "bh"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--bodies", type=int, default=0, help="override N (parity/debug runs; the headline is the default)")
    ap.add_argument("--ratio", type=float, default=10.0, help="Barnes-Hut distance_to_node_radius_ratio")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine tunable name=value (nb200_set_option), repeatable")
    return ap.parse_args()


# ---- inputs ---------------------------------------------------------------------
def make_inputs(n, precision):
    """G1 synthetic galaxy pair. From the compiled reference's make_universe when oracle/_ref travels with
    the repo (exact libstdc++ RNG stream), else a numpy restatement of the same geometry."""
    from oracle import refharness as R
    dtype = np.float64 if precision == "f64" else np.float32
    if R.available(precision):
        lib = R.load(precision)
        d = R.Data(lib).make_universe(n // 2)
        y, m = d.export()
        d.close()
        if m.size == n:
            return y, m, "synthetic: nbody_data::make_universe(%d,100,100,100) via oracle/_ref" % (n // 2)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import numpy_universe
    y, m = numpy_universe(n)
    return y.astype(dtype), m.astype(dtype), "synthetic: numpy restatement of make_universe geometry"


# ---- clocks -----------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline / reference arm -----------------------------------------------------
def cpu_reference_rate(workload, precision, ratio, reps=1, n_direct=None):
    """Reference CPU engine on the host cores, bounded sample. Returns dict(value, unit, cores, kind, sample)."""
    from oracle import refharness as R
    cores = os.cpu_count() or 1
    if workload == "direct":
        n = n_direct or N_CPU_SAMPLE
        y, m, _ = make_inputs(n, precision)
        if R.available(precision):
            lib = R.load(precision)
            cores = lib.nbref_max_threads()
            d = R.Data(lib).import_(y, m)
            e = R.Engine(lib, engine="openmp")
            assert e.init(d)
            sec = e.time_fcompute(reps)
            e.close()
            d.close()
            kind, what = "reference", "nbody_engine_openmp::fcompute"
        else:
            from oracle.oracle import Oracle
            o = Oracle(precision)
            cores = o.threads()
            o.fcompute_openmp(y[:6 * 1024], m[:1024])
            t0 = time.perf_counter()
            o.fcompute_openmp(y, m)
            sec = time.perf_counter() - t0
            kind, what = "port", "oracle/nbody_oracle.c orc_fcompute_openmp"
        return {"value": n * n / sec, "unit": "pair interactions/s", "cores": cores, "kind": kind, "bodies": n, "seconds": sec,
                "sample": "%s, best of %d fcompute at N=%d after one warm-up (%.1f s each); rate is N-independent, "
                          "N=1M extrapolates to %.0f s" % (what, max(1, reps), n, sec, (N_DIRECT ** 2) / (n * n / sec))}
    n = 1 << 18
    y, m, _ = make_inputs(n, precision)
    if R.available(precision):
        lib = R.load(precision)
        cores = lib.nbref_max_threads()
        d = R.Data(lib).import_(y, m)
        e = R.Engine(lib, engine="simple_bh", distance_to_node_radius_ratio=ratio, traverse_type="nested_tree",
                     tree_layout="heap_stackless")
        assert e.init(d)
        sec = e.time_fcompute(0)
        e.close()
        d.close()
        kind, what = "reference", "nbody_engine_simple_bh_heap_stackless::fcompute"
    else:
        from oracle.oracle import Oracle
        o = Oracle(precision)
        cores = o.threads()
        t0 = time.perf_counter()
        tree = o.heap_build(y, m, ratio)
        o.fcompute_bh(y, m, tree)
        sec = time.perf_counter() - t0
        kind, what = "port", "oracle/nbody_oracle.c orc_heap_build + orc_fcompute_bh"
    return {"value": sec * 1e3, "unit": "ms/step", "cores": cores, "kind": kind,
            "sample": "%s, one fcompute (build + walk) at N=%d, ratio %g" % (what, n, ratio)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, timed alone on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    # W warm-up + K timed steps, each step one reference fcompute on a bounded sample (N_CPU_STEP bodies)
    res = cpu_reference_rate(args.workload, args.precision, args.ratio, reps=max(1, args.steps), n_direct=N_CPU_STEP)
    wall = time.perf_counter() - t0
    direct = args.workload == "direct"
    line = {
        "impl": "reference",
        "metric": "pair interactions/s (FP64 direct all-pairs fcompute)" if direct else "Barnes-Hut fcompute ms/step",
        "value": res["value"], "unit": res["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (res["seconds"] * 1e3) if direct else res["value"],
        "higher_is_better": direct, "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": ("direct all-pairs fcompute, reference CPU engine on a bounded sample of the N=1M workload"
                                if direct else "Barnes-Hut heap_stackless fcompute, reference CPU engine, bounded sample"),
                   "sample": res["sample"], "wall_s": wall},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": res["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---- nb200 arm ---------------------------------------------------------------------------
def pinned_array(nbytes, dtype):
    """Page-locked host array via torch (plumbing only)."""
    import torch
    t = torch.empty(nbytes // np.dtype(dtype).itemsize, dtype=torch.float64 if dtype == np.float64 else torch.float32).pin_memory()
    return t, t.numpy()


def run_nb200(args):
    import torch  # noqa: F401  (device selection / pinned memory / rendezvous only)
    from nbody_b200 import Engine, dist, new_unique_id

    _, world_env, local_env = dist.env_rank()
    if world_env > 1:
        torch.cuda.set_device(local_env)          # before the NCCL process group touches a device
    rank, world, local = dist.init_process_group()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE is %d: launch with torch.distributed.run" % (args.gpus, world))
    direct = args.workload == "direct"
    n = args.bodies or (N_DIRECT if direct else N_BH)
    precision = args.precision
    dtype = np.float64 if precision == "f64" else np.float32
    y, m, data_note = make_inputs(n, precision)

    uid = dist.exchange_unique_id(lambda: new_unique_id(precision)) if world > 1 else None
    eng = Engine(precision=precision, devices=[local], rank=rank, nranks=world, uid=uid,
                 kind="direct" if direct else "bh", distance_to_node_radius_ratio=args.ratio,
                 tree_layout="heap_stackless", tree_build_rate=0)
    for item in args.opt:
        k, v = item.split("=")
        eng.set_option(k, int(v))
    if not eng.init(y, m):
        raise SystemExit("engine init failed: " + eng.last_error())
    ybuf = eng.get_y()
    fbuf = eng.create_buffer(ybuf.size())
    flush = eng.create_buffer(256 << 20)          # > 126 MB L2

    def step():
        eng.fill_buffer(flush, 0)                 # L2 flush between timed iterations
        eng.fcompute(0.0, ybuf, fbuf)

    for _ in range(args.warmup):
        step()
    eng.synchronize()

    # ---- timed region: K steps, device time on the engine's stream, max over ranks ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    eng.synchronize()
    launches0 = eng.launch_count()
    eng.mark(0)
    for _ in range(args.steps):
        step()
    eng.mark(1)
    total_ms = eng.elapsed_ms(0, 1)
    eng.synchronize()
    dist.barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = dist.max_over_ranks(total_ms)
    # phase split of the last step (CUDA events recorded inside the library around each kernel)
    phases = eng.last_fcompute_ms()
    force_ms = dist.max_over_ranks(phases["force"])

    # ---- e2e: host buffers through the public API, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        keep_y, host_y = pinned_array(ybuf.size(), dtype)
        keep_f, host_f = pinned_array(ybuf.size(), dtype)
        host_y[:] = y
        for _ in range(1):
            eng.write_from(ybuf, host_y.ctypes.data)
            eng.fcompute(0.0, ybuf, fbuf)
            eng.read_into(host_f.ctypes.data, fbuf)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.write_from(ybuf, host_y.ctypes.data)      # H2D of this step's inputs (pinned)
            eng.fcompute(0.0, ybuf, fbuf)
            eng.read_into(host_f.ctypes.data, fbuf)       # D2H of the step's result
        e2e_s = dist.max_over_ranks(time.perf_counter() - t0)
        e2e = (e2e_s, int(ybuf.size()), int(ybuf.size()), float(np.abs(host_f).max()))

    # ---- Barnes-Hut: exact node-visit / interaction counts of one more (untimed) walk, for the algorithmic bytes ----
    walk_counts = None
    if not direct:
        eng.bh_walk_stats(True)
        eng.fcompute(0.0, ybuf, fbuf)
        eng.synchronize()
        v, k = eng.bh_walk_stats(False)
        walk_counts = (dist.sum_over_ranks(v), dist.sum_over_ranks(k))

    # ---- FP64 FMA peak probe (same process, same clocks) ----
    fma_peak = eng.probe_fma_peak(300.0) if rank == 0 else 0.0

    if rank == 0:
        ms_per_step = total_ms / args.steps
        if direct:
            pairs = float(n) * float(n)
            value = pairs * args.steps / (total_ms * 1e-3)
            unit = "pair interactions/s"
            metric = "pair interactions/s (%s direct all-pairs fcompute, N=%d)" % ("FP64" if precision == "f64" else "FP32", n)
            # dominant kernel: direct_pairs; each rank's launch covers n/world targets x n sources
            pairs_per_launch = pairs / world
            slots, issued = SLOTS_PER_PAIR[precision], ISSUED_PER_PAIR[precision]
            sym_edge = eng.last_direct_path()
            kernel = "direct_pairs"
            if sym_edge:
                # symmetric tiles: 21 FP64-pipe (16 FP32-pipe) instructions per UNORDERED pair = 10.5 (8) per interaction
                issued = 10.5 if precision == "f64" else 4   # FP32: 16 packed two-wide instructions per 2 unordered pairs
                kernel = "%s (tile edge %d)" % ("direct_sym_tiles<4,2>" if precision == "f64" else "direct_sym_tiles_f32x2<8>", sym_edge)
            achieved = pairs_per_launch * 2 * slots / (force_ms * 1e-3) / 1e12
            peak = fma_peak * 2 / 1e12
            roofline = {"bound": "fp64_fma_pipe" if precision == "f64" else "fp32_fma_pipe",
                        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                        "traffic": None, "kernel": kernel, "kernel_ms": force_ms,
                        "algorithmic_per_unit": "%d FMA-pipe slots = %d flop per pair interaction (SURVEY 8d); kernel issues %g per interaction%s"
                                                % (slots, 2 * slots, issued,
                                                   " -- it evaluates each unordered pair once (Newton's third law), so frac by the "
                                                   "ordered-pair convention can exceed 1; frac_issued is the pipe utilisation" if sym_edge else ""),
                        "frac_issued": (pairs_per_launch * issued / (force_ms * 1e-3)) / fma_peak if fma_peak else None,
                        "peak_source": "nb200_probe_fma_peak (FMA chain kernel of this precision, this run); MEASURED_PEAKS.json has no "
                                       "FP64/FP32 vector entry; nominal 148 SM x 64 (FP64) / 128 (FP32) FMA/clk x 1.965 GHz = 37.2 / 74.4 TFLOP/s",
                        "traffic_note": "ordered-pair kernel at N=1M (profiles/r1_ncu_direct_pairs.csv): 48 MB DRAM read + 60 MB written per launch"}
            if sym_edge == 8192 and n == N_DIRECT and world == 1:
                # one ncu --set full capture of this exact launch (profiles/r1_ncu_direct_sym_tiles_n1m.csv):
                # dram__bytes_read.sum 0.314 GB + dram__bytes_write.sum 3.214 GB (the tile partials) per launch
                roofline["traffic"] = 3.528e9
                roofline["traffic_source"] = "profiles/r1_ncu_direct_sym_tiles_n1m.csv (bytes per launch; compute-bound kernel)"
            e2e_obj = None
            if e2e:
                e2e_obj = {"value": pairs * args.steps / e2e[0], "unit": unit, "h2d_bytes_per_step": e2e[1],
                           "d2h_bytes_per_step": e2e[2], "result_maxabs": e2e[3]}
            hib = True
        else:
            value = ms_per_step
            unit = "ms/step"
            metric = "Barnes-Hut heap_stackless fcompute ms/step (N=%d, ratio %g)" % (n, args.ratio)
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
            # bytes = visits*4T + interactions*T + N*(4 + 4T + 3T + 6T) -- what the reference's
            # kfcompute_heap_bh_stackless touches per target; most visits are L1/L2 hits, so this may exceed HBM peak.
            tsz = 8 if precision == "f64" else 4
            visits, inter = walk_counts
            alg_bytes = visits * 4 * tsz + inter * tsz + n * (4 + 13 * tsz)
            achieved = alg_bytes / world / (force_ms * 1e-3) / 1e9
            peak = hbm or 6650.0
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "kernel": "bh_walk_warp_multi<2> (two targets per lane)", "kernel_ms": force_ms, "phases_ms": phases,
                        "node_visits": visits, "interactions": inter, "algorithmic_bytes": alg_bytes,
                        "note": "algorithmic bytes count every per-target node visit; the warp-coherent walk loads each node once per warp "
                                "and the upper tree stays in L1/L2, so the fraction can exceed 1 (SURVEY 8d says so); see profiles/ for DRAM bytes",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm else "fallback 6.65 TB/s (B200_PROFILING.md)"}
            if n == N_BH and world == 1 and args.ratio == 10.0 and precision == "f64":
                # one ncu --set full capture of this exact launch (profiles/r1_ncu_bh_walk_multi_n4m.csv):
                # dram__bytes_read.sum 2.942 GB + dram__bytes_write.sum 0.784 GB per launch; L2 hit 96.5 %, L1 hit 56 %
                roofline["traffic"] = 3.726e9
                roofline["traffic_source"] = "profiles/r1_ncu_bh_walk_multi_n4m.csv (bytes per launch)"
            e2e_obj = None
            if e2e:
                e2e_obj = {"value": e2e[0] * 1e3 / args.steps, "unit": unit, "h2d_bytes_per_step": e2e[1],
                           "d2h_bytes_per_step": e2e[2], "result_maxabs": e2e[3]}
            hib = False
        # context row: the reference's own CUDA kernel recompiled for sm_100a, same box, same inputs (N bounded for direct)
        ref_cuda = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import refcuda
                if refcuda.available(precision):
                    if direct:
                        nr = min(n, 262144)
                        yr, mr, _ = (y, m, None) if nr == n else make_inputs(nr, precision)
                        _, ms_ref = refcuda.direct(yr, mr, block_size=256, reps=1, precision=precision)
                        ref_cuda = {"kernel": "kfcompute + kfcompute_xyz (nbody_engine_cuda_impl.cu:10-124) recompiled for sm_100a, block 256",
                                    "bodies": nr, "value": float(nr) * nr / (ms_ref * 1e-3), "unit": unit, "ms": ms_ref}
                    else:
                        tree = eng.bh_export_tree()
                        _, ms_ref = refcuda.bh_stackless(y, tree[0], tree[1], tree[2], block_size=256, reps=1, precision=precision)
                        ref_cuda = {"kernel": "kfcompute_heap_bh_stackless (nbody_engine_cuda_impl.cu:372-451) recompiled for sm_100a, block 256, "
                                              "walk only on nb200's tree (the reference adds a CPU tree build + transfers per step)",
                                    "bodies": n, "value": ms_ref, "unit": "ms (walk only)"}
            except Exception as exc:
                ref_cuda = {"error": repr(exc)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_rate(args.workload, precision, args.ratio)
            except Exception as exc:  # the baseline must never sink the bench line
                cpu = {"value": None, "unit": unit, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (exc,)}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": hib, "scaling": "strong", "vs_baseline": None,
            "dtype": precision, "data": data_note,
            "config": {"workload": ("direct all-pairs fcompute N=%d (BASELINE configs[2]), targets sharded over %d GPU(s)" % (n, world)) if direct
                       else ("Barnes-Hut heap_stackless fcompute N=%d ratio %g (BASELINE configs[3]), walk sharded over %d GPU(s)" % (n, args.ratio, world)),
                       "bodies": n, "shards": world, "collective": "NCCL all-gather of packed (x,y,z,m) sources per fcompute" if world > 1 else "none",
                       "l2": "256 MiB fill kernel between timed iterations (L2 flush, inside the timed region)",
                       "phases_ms_last_step": phases},
            "clocks": clocks, "e2e": e2e_obj, "gpu_launches": int(launches) * world,
            "roofline": roofline, "cpu_baseline": cpu, "reference_cuda_kernel": ref_cuda,
        }
        print(json.dumps(line))
    eng.free_buffer(fbuf)
    eng.free_buffer(flush)
    eng.close()
    dist.shutdown()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_nb200(args)


if __name__ == "__main__":
    sys.exit(main())
