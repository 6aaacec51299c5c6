"""Data parallelism through the C ABI's own NCCL plane (s2s_dp_init / s2s_dp_allreduce, csrc/dp_nccl.cu) on real GPUs:
two ranks x 16 utterances must reproduce one rank x 32 utterances -- the summed gradient and the parameters after the
replicated gradient step (/B_global, clip, adadelta, row-norm; timit/timit.lua:291-348) -- with the plain all-reduce and
with the bucketed overlap inside s2s_model_fwdbwd, eager and as nodes of the captured / replayed CUDA graph.  Needs >= 2 GPUs (gpurun --gpus 2);
the 128-byte NCCL unique id travels through a gloo store, nothing else does."""
import os
import socket

import numpy as np
import pytest
import torch

from tests.util import make_batch

pytestmark = pytest.mark.gpu

CFG = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=16, KF=10, M=64, MW=7)
B, L, T = 32, 120, 20


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run_rank(rank, world, port, overlap, q):
    import torch.distributed as dist
    import s2s_b200 as s2s
    torch.cuda.set_device(rank)
    store = dist.TCPStore("127.0.0.1", port, world, is_master=(rank == 0))
    ctx = s2s.Context(rank)
    if overlap == "eager":
        ctx.set_graphs(False)      # "graph": the collectives are nodes of the step's captured CUDA graph (call 2 captures, call 3 replays)
    s2s.dp.init(ctx, rank, world, store=store, overlap=bool(overlap))
    assert s2s.dp.nccl_world(ctx) == world
    dev = torch.device("cuda", rank)
    P0 = torch.from_numpy(s2s.init_params(CFG, seed=1234)).to(dev)
    X, lengths, labels, tlens = make_batch(CFG, B, L, T, seed=77)
    lo, hi = s2s.dp.shard_bounds(B, world, rank)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Xs, ls, ys, ts = d(X[lo:hi]), d(lengths[lo:hi]), d(labels[lo:hi]), d(tlens[lo:hi])
    P = P0.clone(); G = torch.zeros_like(P); v = torch.zeros_like(P); a = torch.zeros_like(P)
    outs = []
    for step in range(3):          # eager, captured, replayed
        G.zero_()
        s2s.model_fwdbwd(ctx, CFG, P, G, Xs, ys, lengths=ls, tlens=ts, flags=s2s.NORMALIZE_NLL)
        if not overlap:
            s2s.dp.allreduce(ctx, G)
        torch.cuda.synchronize()
        if step == 0:
            outs.append(G.cpu().numpy().copy())
        s2s.dp.gradient_step(ctx, s2s, CFG, P, G, v, a, B)
    torch.cuda.synchronize()
    outs.append(P.cpu().numpy().copy())
    res = None
    if rank == 0:   # the single-rank big-batch reference on the same GPU, a second context without a communicator
        c1 = s2s.Context(0)
        P1 = P0.clone(); G1 = torch.zeros_like(P1); v1 = torch.zeros_like(P1); a1 = torch.zeros_like(P1)
        g_first = None
        for step in range(3):
            G1.zero_()
            s2s.model_fwdbwd(c1, CFG, P1, G1, d(X), d(labels), lengths=d(lengths), tlens=d(tlens), flags=s2s.NORMALIZE_NLL)
            torch.cuda.synchronize()
            if step == 0:
                g_first = G1.cpu().numpy().copy()
            s2s.dp.gradient_step(c1, s2s, CFG, P1, G1, v1, a1, B)
        torch.cuda.synchronize()
        res = (g_first, P1.cpu().numpy().copy())
        c1.close()
    q.put((rank, outs[0], outs[1], res))
    s2s.dp.destroy(ctx)
    ctx.close()


@pytest.mark.parametrize("overlap", [False, "eager", "graph"])
def test_two_ranks_reproduce_the_single_rank_big_batch(overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_run_rank, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    g_ref, p_ref = res[0][3]
    scale_g = np.abs(g_ref).max()
    for rank, g, p, _ in res:
        assert np.abs(g - g_ref).max() / scale_g < 2e-5, (rank, "summed gradient")          # fp32 sums in a different order
        assert np.abs(p - p_ref).max() / np.abs(p_ref).max() < 2e-5, (rank, "parameters after 3 steps")
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])    # ranks stay bit-identical
