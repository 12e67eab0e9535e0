"""Multi-GPU plumbing: one process per GPU (torch.distributed, backend nccl over NVLink).

* Closed loop: scenarios are independent -> contiguous shards, NO data-path collective.
* EDMD over a sharded snapshot set: each rank accumulates its local Gram pack, ONE all-reduce
  (sum, fp64, ~1.4 kB) makes it global, every rank solves the small systems redundantly and gets
  bitwise-identical A, B, C.
The reduction is backend-agnostic (`gloo` in the CPU tests, `nccl` on the GPU box)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous shard [lo, hi) of `total` items for `rank` of `world` (sizes differ by <= 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_pack(pack, group=None):
    """Sum the Gram pack over ranks in place (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    return pack


def edmd_sharded(local_gram_fn, solve_fn, group=None):
    """local_gram_fn() -> this rank's pack tensor; solve_fn(pack) -> (A, B, C, ...).  The product
    passes the CUDA kernels (edmd.gram_from_snapshots / edmd.edmd_solve)."""
    pack = allreduce_pack(local_gram_fn(), group)
    return solve_fn(pack)


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
