#!/usr/bin/env python
"""nb200 benchmark: BASELINE.json's headline metric on its own configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl nb200|reference] [--workload direct|bh|both]

BASELINE.json's metric has two halves; ONE JSON line carries both (default `--workload both`):

  headline (`metric`, `value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE configs[2] -- direct all-pairs
      fcompute, N = 1,048,576 bodies (G1 synthetic galaxy pair), FP64, target bodies sharded over the N GPUs (one
      process per GPU under torchrun, packed source bodies all-gathered with NCCL inside every fcompute). A "step"
      is one fcompute pass over the whole system: N^2 pair interactions (self pairs included, as the reference's
      block/cuda kernels count them). `value` = N^2 * K / (device time of K steps, max over ranks), inputs resident
      in HBM.
  `bh` block: BASELINE configs[3] -- Barnes-Hut heap_stackless fcompute (tree build + node update + walk) at
      N = 4,194,304, opening ratio 10, walk sharded over the same GPUs; ms/step with the same keys (value, e2e,
      roofline, cpu_baseline, phases).

`e2e`   = the same metric through the public engine API with HOST buffers: every step writes y from pinned host
          memory (write_buffer), runs fcompute and reads f back (each rank its own shard, nb200_read_local).
`roofline` direct: the all-pairs kernel against the FP64 FMA pipe -- SURVEY.md 8(d)'s 18 FP64-pipe instruction
          slots per pair (36 flop) x pairs per launch / that kernel's CUDA-event time; `peak` is a DFMA probe
          kernel timed in the same process (MEASURED_PEAKS.json has no FP64 entry; said in `peak_source`).
          bh: see bh_roofline().
`cpu_baseline` = the reference's own CPU engines (oracle/_ref: nbody_engine_block and nbody_engine_openmp for the
          direct sum, nbody_engine_simple_bh for Barnes-Hut) on ALL host cores (omp_set_num_threads, so torchrun's
          OMP_NUM_THREADS=1 does not apply), on a bounded sample, rank 0 at N = 1 only.

--impl reference times the reference's CPU engine alone (same metric / unit / config.workload) on the bounded sample;
under torchrun rank 0 alone runs it.
"""
import argparse
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
N_CPU_SAMPLE = 131072        # cpu_baseline sample of the direct sum: 1.7e10 pairs, a few seconds per fcompute
N_CPU_STEP = 65536           # --impl reference: bodies per timed step (4.3e9 pairs, < 1 s on 16 cores)
N_CPU_BH = 1 << 18           # simple_bh sample (N = 4M would take minutes per fcompute)
CPU_BUDGET_S = 20.0          # the reference arm stops adding timed steps beyond this
# FMA-pipe instruction slots per pair -- FP64: 18 (the kernel issues 17); FP32: 13 (+1 MUFU on XU)
SLOTS_PER_PAIR = {"f64": 18, "f32": 13}
ISSUED_PER_PAIR = {"f64": 17, "f32": 13}
# Barnes-Hut force evaluation: FMA-pipe instruction slots per accepted (target, node) interaction -- 3 DADD for the
# separation, 3 for d^2, 7 for m r^-3 from the MUFU seed (one series step), three accumulating DFMA
BH_SLOTS_PER_INTERACTION = {"f64": 16, "f32": 12}


def metric_name(workload, precision):
    fp = "FP64" if precision == "f64" else "FP32"
    if workload == "bh":
        return "Barnes-Hut heap_stackless fcompute ms/step (%s) at N=4M" % fp
    return "pair interactions/s (%s direct all-pairs fcompute) at N=1M" % fp


def workload_name(workload, n, ratio):
    if workload == "bh":
        return "Barnes-Hut heap_stackless fcompute N=%d ratio %g, G1 galaxy pair (BASELINE configs[3])" % (n, ratio)
    return "direct all-pairs fcompute N=%d, G1 galaxy pair (BASELINE configs[2])" % n


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nb200", choices=["nb200", "reference"])
    ap.add_argument("--workload", default="both", choices=["direct", "bh", "both"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--bodies", type=int, default=0, help="override N of the headline workload (debug runs)")
    ap.add_argument("--bh-bodies", type=int, default=0, help="override N of the Barnes-Hut block (debug runs)")
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
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _ref_lib(precision):
    """The compiled reference (best build for this host), with its OpenMP team set to every host core: torchrun exports
    OMP_NUM_THREADS=1, which omp_set_num_threads overrides."""
    from oracle import refharness as R
    if not R.available(precision):
        return None, None, 0
    variant = R.host_variant(precision)
    lib = R.load(precision, variant)
    lib.nbref_set_threads(host_cores())
    march = "-march=x86-64-v4 (AVX-512)" if variant == "v4" else "-march=x86-64-v3 (AVX2)"
    return lib, march, lib.nbref_max_threads()


def cpu_direct(precision, n, warmup, steps, budget_s=CPU_BUDGET_S):
    """W untimed + K timed fcompute of the reference's CPU direct-sum engines at `n` bodies; the engine with the
    higher rate is the baseline (both are reported). Timed steps stop early once `budget_s` is spent."""
    from oracle import refharness as R
    y, m, _ = make_inputs(n, precision)
    lib, march, cores = _ref_lib(precision)
    if lib is None:
        from oracle.oracle import Oracle
        o = Oracle(precision)
        o.lib.orc_set_threads(host_cores())
        o.fcompute_openmp(y[:6 * 1024], m[:1024])
        t0 = time.perf_counter()
        o.fcompute_openmp(y, m)
        sec = time.perf_counter() - t0
        return {"value": n * n / sec, "unit": "pair interactions/s", "cores": o.threads(), "kind": "port", "bodies": n,
                "seconds_per_step": sec, "steps_timed": 1,
                "sample": "oracle/nbody_oracle.c orc_fcompute_openmp, one fcompute at N=%d (%.1f s)" % (n, sec)}
    d = R.Data(lib).import_(y, m)
    rates = {}
    done = {}
    for name in ("block", "openmp"):
        e = R.Engine(lib, engine=name)
        assert e.init(d)
        f = e.create_buffer(e.problem_size() * np.dtype(lib.dtype).itemsize)
        for _ in range(max(1, warmup) if name == "block" else 1):
            e.fcompute(0.0, e.get_y(), f)
        k = 0
        t0 = time.perf_counter()
        while k < max(1, steps if name == "block" else min(steps, 2)):
            e.fcompute(0.0, e.get_y(), f)
            k += 1
            if time.perf_counter() - t0 > (budget_s if name == "block" else budget_s / 4):
                break
        sec = (time.perf_counter() - t0) / k
        e.free_buffer(f)
        e.close()
        rates[name] = n * n / sec
        done[name] = (k, sec)
    d.close()
    best = max(rates, key=rates.get)
    k, sec = done[best]
    return {"value": rates[best], "unit": "pair interactions/s", "cores": cores, "kind": "reference", "bodies": n,
            "seconds_per_step": sec, "steps_timed": k, "engine": "nbody_engine_" + best, "build": march,
            "engines": {"nbody_engine_" + kk: vv for kk, vv in rates.items()},
            "sample": "nbody_engine_%s::fcompute (the faster of the reference's block / openmp engines), mean of %d "
                      "fcompute at N=%d after warm-up (%.2f s each); the rate is N-independent, N=1M extrapolates "
                      "to %.0f s per fcompute" % (best, k, n, sec, (N_DIRECT ** 2) / rates[best])}


def cpu_bh(precision, ratio, n=N_CPU_BH):
    """One reference simple_bh fcompute (build + walk) at `n` bodies on all host cores."""
    from oracle import refharness as R
    y, m, _ = make_inputs(n, precision)
    lib, march, cores = _ref_lib(precision)
    if lib is None:
        from oracle.oracle import Oracle
        o = Oracle(precision)
        o.lib.orc_set_threads(host_cores())
        t0 = time.perf_counter()
        tree = o.heap_build(y, m, ratio)
        o.fcompute_bh(y, m, tree)
        sec = time.perf_counter() - t0
        return {"value": sec * 1e3, "unit": "ms/step", "cores": o.threads(), "kind": "port", "bodies": n,
                "sample": "oracle/nbody_oracle.c orc_heap_build + orc_fcompute_bh, one fcompute at N=%d, ratio %g" % (n, ratio)}
    d = R.Data(lib).import_(y, m)
    e = R.Engine(lib, engine="simple_bh", distance_to_node_radius_ratio=ratio, traverse_type="nested_tree",
                 tree_layout="heap_stackless")
    assert e.init(d)
    sec = e.time_fcompute(0)
    e.close()
    d.close()
    return {"value": sec * 1e3, "unit": "ms/step", "cores": cores, "kind": "reference", "bodies": n, "build": march,
            "sample": "nbody_engine_simple_bh_heap_stackless::fcompute, one fcompute (build + walk) at N=%d, ratio %g; "
                      "N=4M takes minutes per fcompute on the host" % (n, ratio)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, timed alone on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    headline = "bh" if args.workload == "bh" else "direct"
    if headline == "direct":
        res = cpu_direct(args.precision, N_CPU_STEP, args.warmup, args.steps)
        value, ms = res["value"], res["seconds_per_step"] * 1e3
    else:
        res = cpu_bh(args.precision, args.ratio)
        value, ms = res["value"], res["value"]
    n = args.bodies or (N_DIRECT if headline == "direct" else N_BH)
    line = {
        "impl": "reference",
        "metric": metric_name(headline, args.precision),
        "value": value, "unit": res["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": headline == "direct", "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": workload_name(headline, n, args.ratio), "bodies": n,
                   "sample": res["sample"], "sample_bodies": res.get("bodies")},
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": res["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload == "both":
        try:
            line["bh"] = {"metric": metric_name("bh", args.precision), "cpu_baseline": cpu_bh(args.precision, args.ratio)}
        except Exception as exc:
            line["bh"] = {"error": repr(exc)}
    line["config"]["wall_s"] = time.perf_counter() - t0
    print(json.dumps(line))
    return 0


# ---- nb200 arm ---------------------------------------------------------------------------
def pinned_array(nbytes, dtype):
    """Page-locked host array via torch (plumbing only)."""
    import torch
    t = torch.empty(nbytes // np.dtype(dtype).itemsize, dtype=torch.float64 if dtype == np.float64 else torch.float32).pin_memory()
    return t, t.numpy()


def measure(args, kind, n, steps, warmup, ctxinfo, want_clocks):
    """One workload on a fresh engine: W warm-up + K timed fcompute (device time, max over ranks), the e2e loop with
    host buffers, Barnes-Hut walk counts. Returns a dict of raw measurements."""
    from nbody_b200 import Engine, dist, new_unique_id
    rank, world, local = ctxinfo
    precision = args.precision
    dtype = np.float64 if precision == "f64" else np.float32
    y, m, data_note = make_inputs(n, precision)
    uid = dist.exchange_unique_id(lambda: new_unique_id(precision)) if world > 1 else None
    eng = Engine(precision=precision, devices=[local], rank=rank, nranks=world, uid=uid, kind=kind,
                 distance_to_node_radius_ratio=args.ratio, tree_layout="heap_stackless", tree_build_rate=0)
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

    for _ in range(warmup):
        step()
    eng.synchronize()

    # ---- timed region: K steps, device time on the engine's stream, max over ranks ----
    sampler = ClockSampler(local)
    if rank == 0 and want_clocks:
        sampler.start()
    dist.barrier()
    eng.synchronize()
    launches0 = eng.launch_count()
    eng.mark(0)
    for _ in range(steps):
        step()
    eng.mark(1)
    total_ms = eng.elapsed_ms(0, 1)
    eng.synchronize()
    dist.barrier()
    out = {"n": n, "steps": steps, "warmup": warmup, "data_note": data_note, "y": y, "m": m}
    out["launches"] = eng.launch_count() - launches0
    out["clocks"] = sampler.stop() if (rank == 0 and want_clocks) else None
    out["total_ms"] = dist.max_over_ranks(total_ms)
    # phase split of the last step (CUDA events recorded inside the library around each kernel)
    out["phases"] = eng.last_fcompute_ms()
    out["force_ms"] = dist.max_over_ranks(out["phases"]["force"])
    out["path"] = eng.last_direct_path() if kind == "direct" else 0

    # ---- e2e: host buffers through the public API, copies inside the timed region; each rank moves its own shard ----
    out["e2e"] = None
    if not args.no_e2e:
        keep_y, host_y = pinned_array(ybuf.size(), dtype)
        keep_f, host_f = pinned_array(ybuf.size(), dtype)
        host_y[:] = y
        host_f[:] = 0
        eng.write_from(ybuf, host_y.ctypes.data)
        eng.fcompute(0.0, ybuf, fbuf)
        eng.read_local_into(host_f.ctypes.data, fbuf)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.write_from(ybuf, host_y.ctypes.data)            # H2D of this step's inputs (pinned): own shard
            eng.fcompute(0.0, ybuf, fbuf)
            eng.read_local_into(host_f.ctypes.data, fbuf)       # D2H of the step's result: own shard
        e2e_s = dist.max_over_ranks(time.perf_counter() - t0)
        # bytes per step summed over ranks: every rank copies 1/world of the state vector each way
        out["e2e"] = (e2e_s, int(ybuf.size()), int(ybuf.size()), dist.max_over_ranks(float(np.abs(host_f).max())))

    # ---- Barnes-Hut: exact node-visit / interaction counts of one more (untimed) walk ----
    if kind == "bh":
        eng.bh_walk_stats(True)
        eng.fcompute(0.0, ybuf, fbuf)
        eng.synchronize()
        v, k = eng.bh_walk_stats(False)
        out["walk_counts"] = (dist.sum_over_ranks(v), dist.sum_over_ranks(k))
        out["walk_profile"] = eng.bh_walk_profile()     # this rank's walk (rank 0 is reported)
    # ---- FP64 / FP32 FMA peak probe (same process, same clocks) ----
    out["fma_peak"] = eng.probe_fma_peak(300.0) if rank == 0 else 0.0
    out["eng"] = eng
    out["bufs"] = (fbuf, flush)
    return out


def finish(out):
    eng = out.pop("eng")
    for b in out.pop("bufs"):
        eng.free_buffer(b)
    eng.close()


def direct_block(args, r, world):
    n, precision = r["n"], args.precision
    pairs = float(n) * float(n)
    value = pairs * r["steps"] / (r["total_ms"] * 1e-3)
    unit = "pair interactions/s"
    force_ms, fma_peak = r["force_ms"], r["fma_peak"]
    pairs_per_launch = pairs / world            # each rank's launch covers 1/world of the pairs
    slots, issued = SLOTS_PER_PAIR[precision], ISSUED_PER_PAIR[precision]
    sym_edge = r["path"]
    kernel = "direct_small" if sym_edge < 0 else "direct_pairs"
    if sym_edge > 0:
        # symmetric tiles: 20 FP64-pipe (16 FP32-pipe) instructions per UNORDERED pair = 10 (8) per interaction
        issued = 10 if precision == "f64" else 4     # FP32: 16 packed two-wide instructions per 2 unordered pairs
        kernel = "%s (tile edge %d)" % ("direct_sym_tiles<8,1>" if precision == "f64" else "direct_sym_tiles_f32x2<8>", sym_edge)
    achieved = pairs_per_launch * 2 * slots / (force_ms * 1e-3) / 1e12
    peak = fma_peak * 2 / 1e12
    roofline = {"bound": "fp64_fma_pipe" if precision == "f64" else "fp32_fma_pipe",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": None,
                "traffic_note": "DRAM bytes cannot be measured in a timed run (ncu serialises and replays); `traffic` is the "
                                "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this very launch, "
                                "committed as profiles/r2_ncu_direct_sym_tiles_n1m.csv, and only filled in when this run is that "
                                "launch (N = 1M, one GPU, FP64, default kernel); it is the tile partials (24 N^2 / T bytes) -- the "
                                "kernel is FP64-pipe bound and DRAM runs at ~5 GB/s",
                "kernel": kernel, "kernel_ms": force_ms,
                "algorithmic_per_unit": "%d FMA-pipe slots = %d flop per pair interaction (SURVEY 8d); kernel issues %g per interaction%s"
                                        % (slots, 2 * slots, issued,
                                           " -- it evaluates each unordered pair once (Newton's third law), so frac by the "
                                           "ordered-pair convention can exceed 1; frac_issued is the pipe utilisation" if sym_edge > 0 else ""),
                "frac_issued": (pairs_per_launch * issued / (force_ms * 1e-3)) / fma_peak if fma_peak else None,
                "peak_source": "nb200_probe_fma_peak (FMA chain kernel of this precision, this run); MEASURED_PEAKS.json has no "
                               "FP64/FP32 vector entry; nominal 148 SM x 64 (FP64) / 128 (FP32) FMA/clk x 1.965 GHz = 37.2 / 74.4 TFLOP/s"}
    if sym_edge == 8192 and n == N_DIRECT and world == 1 and precision == "f64" and not args.opt:
        roofline["traffic"] = 3.537e9      # 0.325 GB read + 3.212 GB written per launch
        roofline["traffic_source"] = "profiles/r2_ncu_direct_sym_tiles_n1m.csv"
    e2e_obj = None
    if r["e2e"]:
        e = r["e2e"]
        e2e_obj = {"value": pairs * r["steps"] / e[0], "unit": unit, "h2d_bytes_per_step": e[1],
                   "d2h_bytes_per_step": e[2], "result_maxabs": e[3],
                   "note": "each rank writes and reads its own shard (nb200_write / nb200_read_local); bytes are summed over ranks"}
    return {"metric": metric_name("direct", precision), "value": value, "unit": unit, "ms_per_step": r["total_ms"] / r["steps"],
            "higher_is_better": True, "roofline": roofline, "e2e": e2e_obj}


def bh_roofline(args, r, world):
    """The walk against the two pipes that bound it (ncu: DRAM traffic is 0.1 % of HBM, so an HBM fraction says
    nothing; SURVEY 8(d)'s algorithmic bytes are kept as extra keys):
      fp64 -- accepted (target, node) interactions x 16 FP64-pipe slots / DFMA peak of this run; the decisions
              themselves are taken in FP32 with a certified margin (FP64 only inside it). The kernel spends these
              slots on (node, 32-target mask) entries, so masked-out lanes lower the fraction: `lane_use` says by
              how much (interactions / (32 x entries)), `frac_issued` is the pipe's share including them;
      hbm  -- SURVEY 8(d)'s bytes per visit / interaction against MEASURED_PEAKS.json hbm_gbs (informational)."""
    precision, n = args.precision, r["n"]
    tsz = 8 if precision == "f64" else 4
    visits, inter = r["walk_counts"]
    force_ms, fma_peak = r["force_ms"], r["fma_peak"]
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks)).get("hbm_gbs") if os.path.exists(peaks) else None
    alg_bytes = visits * 4 * tsz + inter * tsz + n * (4 + 13 * tsz)
    slots = BH_SLOTS_PER_INTERACTION[precision]
    achieved = inter / world * 2 * slots / (force_ms * 1e-3) / 1e12
    peak = fma_peak * 2 / 1e12
    prof = r.get("walk_profile") or {}
    entries = prof.get("entries", 0)
    return {"bound": "fp64_fma_pipe" if precision == "f64" else "fp32_fma_pipe", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "traffic": 3.716e9 if (n == N_BH and world == 1 and precision == "f64" and args.ratio == 10.0 and not args.opt) else None,
            "traffic_note": "`traffic` (when filled in) is dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                            "of this very launch (profiles/r2_ncu_bh_walk_group_n4m.csv: N = 4M, ratio 10, one GPU, FP64): 0.1 % "
                            "of the HBM peak, L2 hit rate 99 % -- the walk is bound by instruction dispatch, not by memory",
            "traffic_source": "profiles/r2_ncu_bh_walk_group_n4m.csv" if (n == N_BH and world == 1 and precision == "f64" and args.ratio == 10.0 and not args.opt) else None,
            "lane_use": (inter / world) / (32.0 * entries) if entries else None,
            "frac_issued": (32.0 * entries * slots / (force_ms * 1e-3)) / fma_peak if (entries and fma_peak) else None,
            "kernel": "Barnes-Hut walk (all walk launches of one fcompute)", "kernel_ms": force_ms, "phases_ms": r["phases"],
            "node_visits": visits, "interactions": inter, "walk_profile_rank0": r.get("walk_profile"),
            "algorithmic_per_unit": "%d FMA-pipe slots = %d flop per accepted (target, node) interaction; node visits are "
                                    "decided on the other pipe and are not counted" % (slots, 2 * slots),
            "peak_source": "nb200_probe_fma_peak (this run)",
            "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / world / (force_ms * 1e-3) / 1e9,
                    "peak_gbs": hbm or 6650.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm else "fallback 6.65 TB/s (B200_PROFILING.md)",
                    "note": "SURVEY 8(d) bytes: visits*4T + interactions*T + N*(4+13T), counted per target; the walk serves "
                            "them from L1/L2 (see profiles/ for dram__bytes and hit rates), so this may exceed the HBM peak"}}


def bh_block(args, r, world):
    unit = "ms/step"
    e2e_obj = None
    if r["e2e"]:
        e = r["e2e"]
        e2e_obj = {"value": e[0] * 1e3 / r["steps"], "unit": unit, "h2d_bytes_per_step": e[1], "d2h_bytes_per_step": e[2],
                   "result_maxabs": e[3]}
    return {"metric": metric_name("bh", args.precision), "value": r["total_ms"] / r["steps"], "unit": unit,
            "ms_per_step": r["total_ms"] / r["steps"], "higher_is_better": False, "steps": r["steps"], "warmup": r["warmup"],
            "config": {"workload": workload_name("bh", r["n"], args.ratio), "bodies": r["n"], "shards": world,
                       "collective": "NCCL all-gather of packed sources + all-gather of leaf-ordered accelerations" if world > 1 else "none",
                       "phases_ms_last_step": r["phases"]},
            "gpu_launches": int(r["launches"]) * world, "roofline": bh_roofline(args, r, world), "e2e": e2e_obj}


def reference_cuda_row(args, kind, r):
    """Context row: the reference's own CUDA kernel recompiled for sm_100a, same box, same inputs (N bounded for direct)."""
    try:
        from oracle import refcuda
        if not refcuda.available(args.precision):
            return None
        if kind == "direct":
            nr = min(r["n"], 262144)
            yr, mr, _ = (r["y"], r["m"], None) if nr == r["n"] else make_inputs(nr, args.precision)
            _, ms_ref = refcuda.direct(yr, mr, block_size=256, reps=1, precision=args.precision)
            return {"kernel": "kfcompute + kfcompute_xyz (nbody_engine_cuda_impl.cu:10-124) recompiled for sm_100a, block 256",
                    "bodies": nr, "value": float(nr) * nr / (ms_ref * 1e-3), "unit": "pair interactions/s", "ms": ms_ref}
        tree = r["eng"].bh_export_tree()
        _, ms_ref = refcuda.bh_stackless(r["y"], tree[0], tree[1], tree[2], block_size=256, reps=1, precision=args.precision)
        return {"kernel": "kfcompute_heap_bh_stackless (nbody_engine_cuda_impl.cu:372-451) recompiled for sm_100a, block 256, "
                          "walk only on nb200's tree (the reference adds a CPU tree build + transfers per step)",
                "bodies": r["n"], "value": ms_ref, "unit": "ms (walk only)"}
    except Exception as exc:
        return {"error": repr(exc)}


def run_nb200(args):
    import torch  # noqa: F401  (device selection / pinned memory / rendezvous only)
    from nbody_b200 import dist

    _, world_env, local_env = dist.env_rank()
    if world_env > 1:
        torch.cuda.set_device(local_env)          # before the NCCL process group touches a device
    rank, world, local = dist.init_process_group()
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE is %d: launch with torch.distributed.run" % (args.gpus, world))
    ctxinfo = (rank, world, local)
    precision = args.precision
    want_cpu = world == 1 and not args.no_cpu_baseline
    headline = "bh" if args.workload == "bh" else "direct"
    line = None
    if headline == "direct":
        n = args.bodies or N_DIRECT
        r = measure(args, "direct", n, args.steps, args.warmup, ctxinfo, True)
        if rank == 0:
            blk = direct_block(args, r, world)
            cpu = None
            if want_cpu:
                try:
                    cpu = cpu_direct(precision, N_CPU_SAMPLE, 1, 2, budget_s=15.0)
                except Exception as exc:  # the baseline must never sink the bench line
                    cpu = {"value": None, "unit": blk["unit"], "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (exc,)}
            line = {
                "metric": blk["metric"], "value": blk["value"], "unit": blk["unit"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": precision, "data": r["data_note"],
                "config": {"workload": workload_name("direct", n, args.ratio), "bodies": n, "shards": world,
                           "collective": "NCCL all-gather of packed (x,y,z,m) sources + reduce-scatter of tile partial sums per fcompute" if world > 1 else "none",
                           "l2": "256 MiB fill kernel between timed iterations (L2 flush, inside the timed region)",
                           "phases_ms_last_step": r["phases"]},
                "clocks": r["clocks"], "e2e": blk["e2e"], "gpu_launches": int(r["launches"]) * world,
                "roofline": blk["roofline"], "cpu_baseline": cpu,
                "reference_cuda_kernel": reference_cuda_row(args, "direct", r) if want_cpu else None,
            }
        finish(r)
    if args.workload in ("bh", "both"):
        n = args.bh_bodies or (args.bodies if headline == "bh" and args.bodies else N_BH)
        if headline == "bh":
            steps, warmup = args.steps, args.warmup
        else:
            steps, warmup = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
        r = measure(args, "bh", n, steps, warmup, ctxinfo, headline == "bh")
        if rank == 0:
            blk = bh_block(args, r, world)
            if want_cpu:
                try:
                    blk["cpu_baseline"] = cpu_bh(precision, args.ratio)
                except Exception as exc:
                    blk["cpu_baseline"] = {"value": None, "unit": "ms/step", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (exc,)}
                blk["reference_cuda_kernel"] = reference_cuda_row(args, "bh", r)
        finish(r)
        if rank == 0 and want_cpu and blk.get("cpu_baseline", {}).get("value"):
            # the same size as the CPU sample on the GPU, so that the two numbers can be divided
            try:
                small = measure(args, "bh", blk["cpu_baseline"]["bodies"], 3, 3, ctxinfo, False)
                blk["cpu_baseline"]["nb200_ms_same_size"] = small["total_ms"] / small["steps"]
                finish(small)
            except Exception as exc:
                blk["cpu_baseline"]["nb200_ms_same_size"] = repr(exc)
        if rank == 0:
            if headline == "bh":
                line = {"metric": blk["metric"], "value": blk["value"], "unit": blk["unit"], "n_gpus": world, "steps": steps,
                        "warmup": warmup, "ms_per_step": blk["ms_per_step"], "higher_is_better": False, "scaling": "strong",
                        "vs_baseline": None, "dtype": precision, "data": r["data_note"], "config": blk["config"],
                        "clocks": r["clocks"], "e2e": blk["e2e"], "gpu_launches": blk["gpu_launches"], "roofline": blk["roofline"],
                        "cpu_baseline": blk.get("cpu_baseline"), "reference_cuda_kernel": blk.get("reference_cuda_kernel")}
                line["config"]["l2"] = "256 MiB fill kernel between timed iterations (L2 flush, inside the timed region)"
            else:
                line["bh"] = blk
    if rank == 0:
        print(json.dumps(line))
    dist.shutdown()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_nb200(args)


if __name__ == "__main__":
    sys.exit(main())
