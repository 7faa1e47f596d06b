"""Host-side multi-GPU logic on CPU: shard arithmetic and the one-process-per-GPU plumbing over
torch.distributed (gloo, world_size 2), as the N > 1 bench path uses it."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_round_trip():
    from nbody_b200 import dist
    n = 48
    y = np.arange(6 * n, dtype=np.float64)
    for g in (1, 2, 4, 8):
        parts = [dist.shard_state(y, n, g, s) for s in range(g)]
        assert all(p.size == 6 * n // g for p in parts)
        assert np.array_equal(dist.unshard_state(parts, n), y)
        lo, hi = dist.shard_range(n, g, g - 1)
        assert hi == n and hi - lo == n // g
    with pytest.raises(ValueError):
        dist.shard_range(10, 4, 0)
    with pytest.raises(ValueError):
        dist.shard_range(8, 4, 4)


def test_sharded_oracle_equals_full(oracle64):
    """Target sharding leaves every body's sum untouched: the per-shard results concatenate to the full f."""
    from nbody_b200 import dist
    g = load_golden_npz("g1_n256")
    n = 256
    full = oracle64.fcompute_openmp(g["y"], g["mass"]).reshape(6, n)
    for shards in (2, 4):
        pieces = []
        for s in range(shards):
            lo, hi = dist.shard_range(n, shards, s)
            pieces.append(oracle64.accel_subset(g["y"], g["mass"], np.arange(lo, hi)))
        assert np.array_equal(np.concatenate(pieces, axis=1), full[3:])


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
from nbody_b200 import dist
rank, world, local = dist.init_process_group(backend="gloo")
assert world == 2
# NCCL unique-id hand-off: rank 0 makes 128 bytes, everyone ends with the same bytes
uid = dist.exchange_unique_id(lambda: bytes(range(128)), 128)
assert uid == bytes(range(128)), uid
# fmaxabs-style reduction and the timing reduction of bench.py
assert dist.max_over_ranks(1.0 + rank) == 2.0
# SPMD shard ownership: each rank owns half of every row; gathering the halves restores the state
n = 16
y = np.arange(6 * n, dtype=np.float64)
mine = dist.shard_state(y, n, world, rank)
import torch, torch.distributed as td
parts = [torch.zeros(6 * n // world, dtype=torch.float64) for _ in range(world)]
td.all_gather(parts, torch.from_numpy(mine))
assert np.array_equal(dist.unshard_state([p.numpy() for p in parts], n), y)
dist.barrier()
print("rank", rank, "ok")
"""


def test_gloo_world_size_2(tmp_path):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % rank in out
