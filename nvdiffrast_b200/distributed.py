"""Data-parallel plumbing for batched renders (one process per GPU, torch.distributed over NCCL / NVLink).

The hot path shards by minibatch image: every kernel is indexed (pixel, image) and no kernel reads another image
(SURVEY.md 8(e)), so there is no data-path collective.  The only exchange is the sum of the gradients of parameters
shared by all images (mesh vertices, texture): one flat all-reduce per step.
"""
import torch


def shard_range(n_total, rank, world):
    """Contiguous [start, stop) slice of a batch of n_total images owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_shared_grads(grads, group=None):
    """Sum the given gradient tensors over all ranks with ONE all-reduce (flatten -> all_reduce -> scatter back in place)."""
    grads = [g for g in grads if g is not None]
    if not grads or not torch.distributed.is_initialized() or torch.distributed.get_world_size(group) == 1:
        return grads
    if len(grads) == 1:
        torch.distributed.all_reduce(grads[0], group=group)
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    torch.distributed.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads
