"""Host-side multi-GPU plumbing (pure torch.distributed, device-agnostic so it is testable with gloo on CPU).

Worlds are independent (no cross-world state in Environment) => contiguous global world ranges per rank,
no data-path collective.  The one exchange: a sum all-reduce of every active brain's [gradient | event count] buffer
between rl_brain_learn and rl_brain_adam, after which all ranks apply the identical Adam step."""
import torch
import torch.distributed as dist


def shard_worlds(n_worlds_global, rank, world_size):
    """-> (n_local, world_id0).  Global world ids are what the RNG is keyed on, so sharding never changes a world."""
    if n_worlds_global % world_size:
        raise ValueError(f"n_worlds={n_worlds_global} must be divisible by the number of ranks ({world_size})")
    n_local = n_worlds_global // world_size
    return n_local, rank * n_local


def allreduce_grads(grads, group=None):
    """One all-reduce (sum) over the concatenation of `grads` (list of 1-D float tensors), written back in place."""
    if not grads:
        return
    if len(grads) == 1:
        dist.all_reduce(grads[0], group=group)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
