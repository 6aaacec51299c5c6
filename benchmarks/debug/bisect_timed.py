"""Bisect helper: full-size (B=32, L=300, T=50) model_fwdbwd vs the float64 oracle for one variant; prints per-output errors.
usage: python benchmarks/debug/bisect_timed.py K ragged lam [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import s2s_b200 as s2s
from oracle.oracle import Oracle, init_params, segment_names, build
from tests.util import make_batch, rel_err, dev
build()
K, ragged, lam = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
L, T = 300, 50
orc = Oracle("f64")
cfg = dict(D=123, H=256, NL=3, S=512, ST=256, V=62, K=K, KF=10, M=64, MW=7)
P = init_params(cfg, seed=1234, dtype=np.float64, oracle=orc)
X, lengths, labels, tlens = make_batch(cfg, B, L, T, seed=1000, ragged=bool(ragged))
ref = orc.model_fwdbwd(cfg, P, X, lengths, labels, tlens, lam=lam, normalize_nll=True, nthreads=len(os.sched_getaffinity(0)))
ctx = s2s.Context(0)
Pd = dev(P, torch.float32); G = torch.zeros_like(Pd)
logp = ctx.new(B, T, cfg["V"]); dX = ctx.new(B, L, cfg["D"]); nll = ctx.new(B)
s2s.model_fwdbwd(ctx, cfg, Pd, G, dev(X), dev(labels), lengths=dev(lengths), tlens=dev(tlens), lam=lam, flags=s2s.NORMALIZE_NLL, nll=nll, logp=logp, dX=dX)
torch.cuda.synchronize()
annot = s2s.model_annotations(ctx, B, L, 512).cpu().numpy()
lp = logp.cpu().numpy(); dx = dX.cpu().numpy(); Gh = G.cpu().numpy()
e_an = [rel_err(annot[b, :lengths[b]], ref["annot"][b, :lengths[b]]) for b in range(B)]
e_lp = [rel_err(lp[b, :tlens[b]], ref["logp"][b, :tlens[b]]) for b in range(B)]
e_dx = [rel_err(dx[b, :lengths[b]], ref["dX"][b, :lengths[b]]) for b in range(B)]
print(f"K={K} ragged={ragged} lam={lam} B={B} env={ {k: v for k, v in os.environ.items() if k.startswith('S2S_')} }")
print(f"  nll {rel_err(nll.cpu().numpy(), ref['nll']):.2e}  annot max {max(e_an):.2e}  logp max {max(e_lp):.2e}  dX max {max(e_dx):.2e} (b={int(np.argmax(e_dx))}, L_b={lengths[int(np.argmax(e_dx))]}, T_b={tlens[int(np.argmax(e_dx))]})")
print("  dX per utt:", " ".join(f"{e:.1e}" for e in e_dx))
print("  lengths:", list(lengths)); print("  tlens:", list(tlens))
segs = {}
for (off, rows, cols), name in zip(orc.param_segments(cfg), segment_names(cfg)):
    a, b_ = Gh[off:off + rows * cols], ref["G"][off:off + rows * cols]
    segs[name] = float(np.abs(a - b_).max() / max(np.abs(b_).max(), 1e-6 * np.abs(ref["G"]).max()))
worst = sorted(segs.items(), key=lambda kv: -kv[1])[:6]
print("  worst grad segments:", ", ".join(f"{k} {v:.1e}" for k, v in worst))
