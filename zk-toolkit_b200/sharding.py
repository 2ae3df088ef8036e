"""Multi-GPU plumbing: one process per GPU, each owns a contiguous shard of the (point, scalar)
vectors, reduces it to ONE partial point on its device, and the partials are exchanged with a
single all-gather (48 words per rank for G1) over NCCL/NVLink (gloo in the CPU tests).  The sum of
partials does not depend on the partition (exact group arithmetic), so the result is bit-identical
for any world size.  Replaces the running `sum` of polynomial.rs:276-280 across devices."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous, balanced: ranks get floor/ceil(n / world) consecutive terms"""
    return n * rank // world, n * (rank + 1) // world


def gather_partials(partial, group=None):
    """partial: 1-D int32 tensor (device tensor under NCCL, CPU tensor under gloo).
    Returns a (world, words) tensor on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partial.reshape(1, -1)
    flat = partial.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.reshape(world, flat.numel())
