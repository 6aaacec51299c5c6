"""Data-parallel host logic on CPU: two gloo ranks each compute the gradient of their shard of a minibatch
(with the CPU oracle standing in for the CUDA engine), all-reduce it through the package's dp helpers, and
must reproduce the single-process full-batch gradient and per-utterance losses (timit/timit.lua:240-295)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import make_batch

SMALL = dict(D=5, H=4, NL=2, S=6, ST=5, V=7, K=2, KF=4, M=3, MW=2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import s2s_b200 as s2s
    from oracle.oracle import Oracle, init_params
    orc = Oracle("f64")
    P = init_params(SMALL, seed=7, dtype=np.float64, oracle=orc)
    X, lengths, labels, tlens = make_batch(SMALL, 5, 9, 4, seed=3)
    lo, hi = s2s.dp.shard_bounds(5, world, rank)
    out = orc.model_fwdbwd(SMALL, P, X[lo:hi], lengths[lo:hi], labels[lo:hi], tlens[lo:hi], want=())
    G = torch.from_numpy(out["G"].copy())
    nll = torch.tensor([out["nll"].sum()])
    assert s2s.dp.world() == (rank, world)
    s2s.dp.allreduce_gradients(G, nll)
    q.put((rank, lo, hi, G.numpy(), float(nll)))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(orc64):
    from oracle.oracle import init_params
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    P = init_params(SMALL, seed=7, dtype=np.float64, oracle=orc64)
    X, lengths, labels, tlens = make_batch(SMALL, 5, 9, 4, seed=3)
    full = orc64.model_fwdbwd(SMALL, P, X, lengths, labels, tlens, want=())
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]           # contiguous shards, sizes differ by <= 1
    for _, _, _, G, nll in res:
        assert np.allclose(G, full["G"], rtol=1e-10, atol=1e-12)     # every rank holds the full-batch gradient
        assert abs(nll - full["nll"].sum()) < 1e-9


def test_shard_bounds():
    import s2s_b200 as s2s
    for n in (1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            b = [s2s.dp.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
