"""Multi-process (one rank per GPU, NCCL) parity check, launched by torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mp_check.py

Every rank holds a body shard; fcompute all-gathers packed sources with NCCL (direct) or replicates the tree and
shards the walk (Barnes-Hut); fmaxabs all-reduces one scalar; read_buffer gathers the shards. Results must equal
the single-GPU engine bit for bit (same per-target summation order) and the CPU oracle to 1e-12."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NBREF_QUIET", "1")


def main():
    import torch
    from nbody_b200 import Engine, dist, new_unique_id
    from util import universe
    from conftest import rel_err_per_body
    from oracle.oracle import Oracle

    rank, world, local = dist.init_process_group()
    torch.cuda.set_device(local)
    n = 65536
    y, m = universe(n)
    orc = Oracle("f64")
    results = {}
    for label, kind, kw, opts in (("direct", "direct", {}, (("direct_symmetric", 0),)),      # ordered pairs + all-gather
                                  ("direct-symmetric", "direct", {}, (("direct_symmetric", 1), ("direct_sym_tile", 2048))),
                                  ("bh", "bh", dict(distance_to_node_radius_ratio=3.1623), ())):
        uid = dist.exchange_unique_id(lambda: new_unique_id("f64"))
        e = Engine(devices=[local], rank=rank, nranks=world, uid=uid, kind=kind, **kw)
        for k, v in opts:
            e.set_option(k, v)
        assert e.shards() == (world, rank)
        assert e.init(y, m), e.last_error()
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0.0, e.get_y(), f)
        got = e.read_buffer(f)                      # collective: every rank receives the full vector
        # shard-local read: only this rank's columns land in the full-layout host array, nothing else is touched
        loc = np.full(got.shape, np.nan)
        e.read_local_into(loc.ctypes.data, f)
        lo, hi = dist.shard_range(n, world, rank)
        mine = np.zeros((6, n), dtype=bool)
        mine[:, lo:hi] = True
        assert np.array_equal(loc.reshape(6, n)[mine], got.reshape(6, n)[mine]) and np.isnan(loc.reshape(6, n)[~mine]).all()
        tmp = e.create_buffer(e.get_y().size())
        e.fmaddn(tmp, e.get_y(), [f], np.array([1e-3]))
        stage = e.read_buffer(tmp)
        # body arrays <-> state vector across ranks: every rank uploads its own body range, every rank reads all bodies
        pos, vel = e.get_bodies()
        assert np.array_equal(pos.T.ravel(), y[:3 * n]) and np.array_equal(vel.T.ravel(), y[3 * n:])
        ps, vs = e.get_bodies(y=tmp)
        assert np.array_equal(np.concatenate([ps.T.ravel(), vs.T.ravel()]), stage)
        assert e.set_option("step_graph", 1) == 0 and e.step_graph_stats()["state"] == "off"   # ignored with ranks
        mx = e.fmaxabs(tmp)                         # NCCL max all-reduce
        assert mx == np.abs(stage).max(), (mx, np.abs(stage).max())
        assert np.array_equal(stage, orc.fmaddn(stage, y, [got], np.array([1e-3])))
        e.close()
        if label == "direct":
            uid = dist.exchange_unique_id(lambda: new_unique_id("f64"))
            with Engine(devices=[local], rank=rank, nranks=world, uid=uid) as eb:
                yr = y.reshape(6, n)
                assert eb.init_bodies(np.ascontiguousarray(yr[0:3].T), np.ascontiguousarray(yr[3:6].T), m)
                assert np.array_equal(eb.read_buffer(eb.get_y()), y)
        results[kind] = got
        if rank == 0:
            with Engine(devices=[local], kind=kind, **kw) as single:
                # same tile/segment shape as a shard would pick is not guaranteed, so compare to tolerance
                assert single.init(y, m)
                fs = single.create_buffer(single.get_y().size())
                single.fcompute(0.0, single.get_y(), fs)
                ref = single.read_buffer(fs)
            err = rel_err_per_body(got, ref, n)
            assert err <= (1e-13 if kind == "direct" else 0.0), err
            if kind == "direct":
                t = np.unique(np.random.RandomState(1).randint(0, n, 128))
                want = orc.accel_subset(y, m, t)
                assert rel_err_per_body(got.reshape(6, n)[3:, t], want, t.size) <= 1e-12
            else:
                tree = orc.heap_build(y, m, 3.1623)
                want, _, _ = orc.fcompute_bh(y, m, tree)
                assert rel_err_per_body(got, want, n) <= 1e-12
            print("mp_check %s ok: %d ranks, vs single-GPU rel err %.2e" % (label, world, err), flush=True)
    dist.barrier()
    print("rank %d done" % rank, flush=True)


if __name__ == "__main__":
    main()
