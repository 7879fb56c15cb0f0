"""Host-side plumbing of the multi-GPU modes (one process per GPU, torch.distributed).

* `shard_slots`      keyframe slot s lives on rank s % world (the keyframe deque is FIFO with eviction
                     of the oldest, odometry/pipeline.cpp:253-257, so round-robin stays balanced).
* `exchange_handles` all-gathers the 64-byte CUDA IPC handles of the ranks' mailboxes (works on any
                     backend: the payload is a byte tensor).
* `connect_peers`    export + exchange + `madicp_comm_connect` for a Registrar.
* `sum_in_rank_order` the reduction rule the persistent kernel applies to the 48-value H/b tiles
                     (every rank adds the partials in rank order => identical bits everywhere); used by
                     the host-driven NCCL baseline and by the CPU (gloo) tests of the sharding logic.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_slots(num_keyframes, rank, world):
    return [s for s in range(num_keyframes) if s % world == rank]


def exchange_handles(handle: bytes, device="cpu", group=None):
    world = dist.get_world_size(group)
    h = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    out = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(out, h, group=group)
    return [bytes(t.cpu().tolist()) for t in out]


def connect_peers(registrar, device, group=None):
    """Wire a Registrar's mailbox to its peers.  Call once, with the GPU idle; afterwards never enqueue
    a collective while cross-GPU registrations are in flight."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = exchange_handles(registrar.comm_export(), device=device, group=group)
    registrar.comm_connect(rank, world, handles)
    dist.barrier(group)
    return rank, world


def sum_in_rank_order(tile, group=None):
    """All ranks receive every rank's tile and add them in rank order (NOT an all_reduce, whose
    association order is backend-defined).  tile: 1-D float64 array/tensor; returns a numpy array."""
    t = torch.as_tensor(np.asarray(tile, dtype=np.float64)).clone()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    acc = parts[0].clone()
    for r in range(1, world):
        acc += parts[r]
    return acc.numpy()
