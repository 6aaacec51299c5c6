"""Batch-sharded data parallelism: the only axis of the path that shards.

Every rank holds the full parameters and optimiser state, runs forward+backward on its shard of the
minibatch, then ONE all-reduce (sum) of the flat gradient makes the gradient step replicated and
deterministic: g /= B_global, clip, adadelta, row-norm (timit/timit.lua:291-348).  torch.distributed is the
plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests); there is no other collective on the path.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, world_size, rank):
    """Contiguous shard [lo, hi) of n_items utterances for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(G, nll=None):
    """Sum the flat gradient (and optionally the per-rank NLL sum) over all ranks, in place."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM)
        if nll is not None:
            dist.all_reduce(nll, op=dist.ReduceOp.SUM)
    return G


def gradient_step(ctx, ops, cfg, P, G, v, a, global_batch, maxnorm=1e20, wd=0.0, rho=0.95, eps=1e-8, colnorm=1.0):
    """The replicated part of the step, after the all-reduce (timit.lua:292-348)."""
    ops.grad_finalize(ctx, G, P, global_batch, maxnorm, wd=wd, want_norm=False)
    ops.adadelta(ctx, P, G, v, a, rho=rho, eps=eps)
    if colnorm:
        ops.model_rownorm_constraint(ctx, cfg, P, colnorm)
