"""Batch-sharded data parallelism: the only axis of the path that shards.

Every rank holds the full parameters and optimiser state, runs forward+backward on its shard of the
minibatch, then the flat gradient is summed over the ranks and the gradient step is replicated and
deterministic: g /= B_global, clip, adadelta, row-norm (timit/timit.lua:291-348).

The collective lives BEHIND THE C ABI (s2s_dp_init / s2s_dp_allreduce on libnccl directly, csrc/dp_nccl.cu), so
a Lua host can use it through FFI; with `overlap` the library reduces the gradient buckets itself on a side
stream under the remaining backward pass.  This module only ships the 128-byte NCCL unique id between the
ranks (a torch.distributed store / process group when one exists -- plumbing, not the data path) and keeps a
`torch.distributed` all-reduce for the CPU (gloo) tests of the host logic.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, world_size, rank):
    """Contiguous shard [lo, hi) of n_items utterances for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---- NCCL plane behind the C ABI -------------------------------------------------------------------------
def _prefer_bundled_nccl():
    """The library dlopens libnccl.so.2 itself ($S2S_NCCL_LIB, a copy already mapped into the process, the system one).  Under a
    Python host, point it at the NCCL wheel torch was built against so that one NCCL version serves the whole process."""
    if os.environ.get("S2S_NCCL_LIB"):
        return
    try:
        import nvidia.nccl as _n
        cand = os.path.join(os.path.dirname(_n.__file__ or list(_n.__path__)[0]), "lib", "libnccl.so.2")
    except Exception:
        cand = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so.2")
    if os.path.exists(cand):
        os.environ["S2S_NCCL_LIB"] = cand


def nccl_available():
    _prefer_bundled_nccl()
    return bool(_lib.load().s2s_dp_available())


def unique_id():
    """ncclGetUniqueId through the C ABI: 128 bytes (call on rank 0, ship to every rank)."""
    _prefer_bundled_nccl()
    buf = (C.c_char * 128)()
    check(_lib.load().s2s_dp_unique_id(buf))
    return bytes(buf.raw)


def init(ctx, rank, world_size, uid=None, store=None, overlap=False):
    """Create the context's NCCL communicator.  The unique id comes from `uid` (bytes), or is exchanged through
    `store` (a torch.distributed Store), or through the initialised torch.distributed process group."""
    if world_size <= 1:
        return
    _prefer_bundled_nccl()
    if uid is None:
        if store is not None:
            if rank == 0:
                uid = unique_id()
                store.set("s2s_nccl_uid", uid)
            else:
                uid = bytes(store.get("s2s_nccl_uid"))
        elif dist.is_available() and dist.is_initialized():
            obj = [unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            uid = obj[0]
        else:
            raise _lib.S2SError("dp.init: need the NCCL unique id (uid=), a store, or an initialised torch.distributed group")
    assert len(uid) == 128
    buf = (C.c_char * 128).from_buffer_copy(uid)
    check(ctx.lib.s2s_dp_init(ctx.h, int(rank), int(world_size), buf))
    if overlap:
        check(ctx.lib.s2s_dp_set_overlap(ctx.h, 1))


def set_overlap(ctx, enable):
    check(ctx.lib.s2s_dp_set_overlap(ctx.h, int(bool(enable))))


def nccl_world(ctx):
    return int(ctx.lib.s2s_dp_world(ctx.h))


def allreduce(ctx, G):
    """G := sum over ranks, in place, on the context's stream (s2s_dp_allreduce)."""
    assert G.is_cuda and G.dtype == torch.float32 and G.is_contiguous()
    check(ctx.lib.s2s_dp_allreduce(ctx.h, C.c_void_p(G.data_ptr()), G.numel()))
    return G


def broadcast(ctx, P, root=0):
    assert P.is_cuda and P.dtype == torch.float32 and P.is_contiguous()
    check(ctx.lib.s2s_dp_broadcast(ctx.h, C.c_void_p(P.data_ptr()), P.numel(), int(root)))
    return P


def destroy(ctx):
    check(ctx.lib.s2s_dp_destroy(ctx.h))


# ---- torch.distributed fallback for the CPU (gloo) tests of the host logic ---------------------------------
def allreduce_gradients(G, nll=None, ctx=None):
    """Sum the flat gradient (and optionally the per-rank NLL sum) over all ranks, in place.  With a context that owns
    an NCCL communicator the C-ABI collective is used; otherwise torch.distributed (gloo in the CPU tests)."""
    if ctx is not None and nccl_world(ctx) > 1:
        allreduce(ctx, G)
        if nll is not None:
            allreduce(ctx, nll)
        return G
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
