"""One process per GPU: rendezvous and shard arithmetic.

torch.distributed is used for plumbing only (rank discovery, a byte broadcast of
the NCCL unique id, barriers and the max-over-ranks reduction of timings). The
data path -- the all-gather of packed source bodies inside fcompute and the
scalar max all-reduce of fmaxabs -- is issued by the native library on its own
NCCL communicator (nbody_b200/csrc/nb200_api.cu).

Sharding rule (include/nb200.h): shard g of G owns bodies [g*N/G, (g+1)*N/G) of
every row of a state vector [rx|ry|rz|vx|vy|vz].
"""
import os

import numpy as np


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n, nshards, shard):
    """Half-open body range owned by `shard`. N must divide evenly (NCCL all-gather needs equal counts)."""
    if nshards < 1 or not 0 <= shard < nshards:
        raise ValueError("bad shard %d of %d" % (shard, nshards))
    if n % nshards != 0:
        raise ValueError("N = %d is not a multiple of the shard count %d" % (n, nshards))
    per = n // nshards
    return shard * per, (shard + 1) * per


def shard_state(y, n, nshards, shard):
    """Columns of `shard` from a full state vector: (6, N) -> flat 6 * N/G."""
    lo, hi = shard_range(n, nshards, shard)
    return np.ascontiguousarray(np.asarray(y).reshape(6, n)[:, lo:hi]).reshape(-1)


def unshard_state(parts, n):
    """Inverse of shard_state for the list of all shards in order."""
    g = len(parts)
    per = n // g
    return np.concatenate([np.asarray(p).reshape(6, per) for p in parts], axis=1).reshape(-1)


def init_process_group(backend=None):
    """Join the torchrun rendezvous; returns (rank, world, local_rank). No-op for world_size 1."""
    rank, world, local = env_rank()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                import torch
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload, nbytes, src=0):
    """Broadcast a fixed-size byte string from `src` (payload may be None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def exchange_unique_id(make_id, nbytes=128):
    """Rank 0 calls make_id() (nb200_comm_unique_id); everyone returns the same bytes."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    uid = make_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(uid, nbytes, src=0)


def max_over_ranks(value):
    """Max of a host scalar over ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def sum_over_ranks(value):
    """Sum of a host integer over ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return int(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def shutdown():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
