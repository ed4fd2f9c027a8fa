"""TEST SCAFFOLDING (the GPU path shards inside workloads.build_c5 and reduces rho in csrc/comm.cu).
Index-slice sharding of the particles over ranks: rank r owns rows
[lo, hi) of every species, the grid is replicated, rho is summed with one all-reduce per step.
No spatial decomposition, hence no particle migration and no halo exchange."""


def slice_for_rank(n, rank, world):
    """Contiguous, balanced (sizes differ by at most one) partition of n rows."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, dist=None):
    """Timing rule of the bench contract: the slowest rank defines the step time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(array, dist=None):
    """In-place sum of a numpy float64 array over ranks (the rho exchange, gloo/NCCL agnostic)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return array
    import torch
    t = torch.from_numpy(array)
    if dist.get_backend() == "nccl":
        tc = t.cuda()
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
        t.copy_(tc.cpu())
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return array
