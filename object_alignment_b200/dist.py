"""Multi-GPU plumbing: one process per GPU (torchrun), base vertices sharded across ranks.

torch.distributed is used ONLY as a rendezvous (gloo: barrier, broadcast of the NCCL unique id,
max-over-ranks of timings).  The data path never touches torch: each rank's engine owns a raw
NCCL communicator inside libicp_b200.so and resolves the global nearest neighbour with one
`ncclAllReduce(MIN, uint64)` over the per-query keys plus one `ncclAllReduce(SUM, double[20])`
over the covariance moments per iteration.
"""
from __future__ import annotations

import os

import numpy as np

from .engine import IcpEngine, shard_range


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))))


def init_rendezvous(backend="gloo"):
    """Join the torchrun rendezvous (MASTER_ADDR/MASTER_PORT from the environment)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        rank, world, _ = env_rank_world()
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def connect_engine(engine: IcpEngine, dist=None):
    """Collective: give every rank's engine the same NCCL unique id and build the communicator."""
    world = engine.params.world_size
    if world == 1:
        return
    dist = dist or init_rendezvous()
    box = [IcpEngine.comm_unique_id() if engine.params.rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.comm_init(box[0])
    dist.barrier()


def allreduce_min_keys_host(keys: np.ndarray, dist) -> np.ndarray:
    """Host-side statement of collective 1 (used by the gloo CPU tests): element-wise MIN of
    uint64 keys.  d2 >= 0, so the keys are < 2^63 and order the same as int64."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(keys).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.numpy().view(np.uint64)


def allreduce_sum_host(values: np.ndarray, dist) -> np.ndarray:
    import torch

    t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def allreduce_max_scalar(x: float, dist) -> float:
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


__all__ = ["env_rank_world", "init_rendezvous", "connect_engine", "allreduce_min_keys_host",
           "allreduce_sum_host", "allreduce_max_scalar", "shard_range"]
