"""Ensemble (data-parallel) sharding of the rollout: one process per GPU, no collective on the data path.

Mirrors the data-parallel surface of ``fme.core.distributed`` that inference uses
(``/root/reference/fme/core/distributed/torch_distributed.py:36-154``):

* ``init_from_env``      -- ``TorchDistributed.__init__`` (:47-78): RANK / WORLD_SIZE / LOCAL_RANK from the
  environment (torchrun), NCCL on GPUs, Gloo on CPU;
* ``member_slice``       -- ``get_local_slices`` / ``local_batch_size`` (:112-128): contiguous, even split of the
  batch (ensemble) dimension; a remainder is rejected exactly like the reference's divisibility check;
* ``gather_members``     -- ``gather`` / the ``reduce_mean`` of the inference aggregators
  (``fme/ace/aggregator/inference/reduced.py:51``): the ONE collective of the rollout, an ``all_gather`` of
  per-member diagnostics ``[B_local, n]`` -> ``[B_global, n]`` in rank order.

The forward pass itself never communicates: every member is an independent trajectory.
"""
import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """(rank, world, local_rank); initialises the default process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = device if device is not None else torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def member_slice(global_members: int, rank: int, world: int) -> slice:
    """Members ``[rank * B, (rank + 1) * B)`` of an ensemble of ``global_members`` (B = global_members / world)."""
    if global_members % world != 0:
        raise ValueError(f"ensemble size {global_members} is not divisible by the number of ranks {world}")
    b = global_members // world
    return slice(rank * b, (rank + 1) * b)


def gather_members(local: torch.Tensor) -> torch.Tensor:
    """all_gather along dim 0 in rank order; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction of the benchmark: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device if device is not None else "cpu", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
