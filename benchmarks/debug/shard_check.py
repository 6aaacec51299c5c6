"""Single-GPU reproduction of the two shards the 2-GPU test feeds its ranks (B=16 each of a ragged 32-utterance batch, L=120, T=20, K=16):
model_fwdbwd + synchronize, compared with the float64 oracle.  usage: python benchmarks/debug/shard_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import s2s_b200 as s2s
from oracle.oracle import Oracle, init_params, build
from tests.util import make_batch, rel_err
build()
orc = Oracle("f64")
CFG = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=16, KF=10, M=64, MW=7)
B, L, T = 32, 120, 20
X, lengths, labels, tlens = make_batch(CFG, B, L, T, seed=77)
P = init_params(CFG, seed=1234, dtype=np.float64, oracle=orc)
ctx = s2s.Context(0)
d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dt or torch.from_numpy(np.ascontiguousarray(a)).dtype).cuda()
Pd = d(P, torch.float32)
for lo, hi in ((0, 16), (16, 32), (0, 32)):
    G = torch.zeros_like(Pd)
    nll = s2s.model_fwdbwd(ctx, CFG, Pd, G, d(X[lo:hi]), d(labels[lo:hi]), lengths=d(lengths[lo:hi]), tlens=d(tlens[lo:hi]), flags=s2s.NORMALIZE_NLL)
    torch.cuda.synchronize()
    ref = orc.model_fwdbwd(CFG, P, X[lo:hi], lengths[lo:hi], labels[lo:hi], tlens[lo:hi], normalize_nll=True, nthreads=16, want=())
    print(f"shard [{lo},{hi}): max len {lengths[lo:hi].max()} max tlen {tlens[lo:hi].max()}  nll err {rel_err(nll.cpu().numpy(), ref['nll']):.2e}  grad err {rel_err(G.cpu().numpy(), ref['G']):.2e}", flush=True)
