"""nb200: B200-native compute engine for drons/nbody (direct + Barnes-Hut fcompute
and the solver state-vector ops) behind the reference's nbody_engine API.

    nbody_b200.engine   Python host mirror of nbody_engine over the C ABI (ctypes)
    nbody_b200.dist     one-process-per-GPU plumbing (torch.distributed rendezvous,
                        NCCL unique-id hand-off, body-shard arithmetic)
    nbody_b200.build    nvcc / g++ build recipe for the in-tree libraries
    nbody_b200/csrc     sm_100a kernels + C ABI (include/nb200.h)
    nbody_b200/host     C++ adapter class nbody_engine_b200 for the reference tree
"""
from .engine import Engine, Memory, NativeLibraryMissing, load_library, device_count, parse_devices, new_unique_id  # noqa: F401

__all__ = ["Engine", "Memory", "NativeLibraryMissing", "load_library", "device_count", "parse_devices", "new_unique_id"]
