"""World sharding across GPUs (no reference counterpart: the reference is
single-GPU, include/madrona/mw_gpu.hpp:122).

Worlds are independent, so GPU g simply owns worlds [g*W, (g+1)*W) with its own
executor; nothing is exchanged inside a step.  The only collective is the
gather of exported tensors (observations / rewards / dones) into one
world-major tensor, and the reverse slice for actions.  Pure torch.distributed
plumbing: works with NCCL on GPUs and with gloo on CPU (tests).
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_range(total_worlds: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first_world, num_worlds) of this rank; shards must be equal so the
    gather is a plain all_gather."""
    if total_worlds % world_size != 0:
        raise ValueError(f"{total_worlds} worlds do not split evenly over {world_size} ranks")
    per = total_worlds // world_size
    return rank * per, per


def world_seed(base_seed: int, first_world: int, local_world: int) -> int:
    """Seed of a world depends only on its GLOBAL index, so an N-GPU run
    simulates exactly the worlds a 1-GPU run of the same total would."""
    return base_seed + first_world + local_world


def gather_exported(local, out=None, group=None):
    """all_gather a per-rank exported tensor [W_local, ...] into [W_total, ...]
    (rank-major == world-major because shards are contiguous world ranges)."""
    import torch
    import torch.distributed as dist

    world_size = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world_size * local.shape[0],) + tuple(local.shape[1:]),
                          dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def local_slice(global_tensor, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
    """The rows of a world-major global tensor (e.g. actions) this rank owns."""
    import torch.distributed as dist

    if rank is None:
        rank = dist.get_rank(group)
    if world_size is None:
        world_size = dist.get_world_size(group)
    first, count = shard_range(global_tensor.shape[0], world_size, rank)
    return global_tensor[first:first + count]


def connect_peer_gather(gather, group=None):
    """Exchange the cudaIpc handles of a madrona_b200.PeerGather over the process group
    (host-side plumbing, works with any backend) and map every peer's buffer."""
    import torch.distributed as dist

    world_size = dist.get_world_size(group)
    handles = [None] * world_size
    dist.all_gather_object(handles, gather.local_handle(), group=group)
    gather.connect(handles)
    dist.barrier(group=group)
    return gather
