"""Decoder (nn.Attention) forward / backward time at the cfg2 shape: B=32, L=300, T=50, ST=256, S=A=512, content attention.
usage: [S2S_DEC_CLUSTER=0] [S2S_DEC_PROF=1] python benchmarks/dec_micro.py [B L T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import s2s_b200 as s2s
B, L, T = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 300, 50)
cfg = dict(D=123, H=256, NL=0, S=512, ST=256, V=62, K=0, KF=4, M=64, MW=7)
ctx = s2s.Context(0)
n = s2s.param_count(cfg)
g = torch.Generator(device="cuda").manual_seed(0)
P = (torch.rand(n, device="cuda", generator=g) - 0.5) * 0.1
h = torch.randn(B, L, 512, device="cuda", generator=g) * 0.5
y = torch.randint(0, 61, (B, T), device="cuda", dtype=torch.int32, generator=g)   # seeded: runs of the two paths are comparable
G = torch.zeros_like(P); dlogp = torch.randn(B, T, 62, device="cuda", generator=g)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
N = 10
for it in range(3 + N):
    ev[0].record(torch.cuda.current_stream())
    logp = s2s.attention_forward(ctx, cfg, P, h, y)
    ev[1].record(torch.cuda.current_stream())
    dh = s2s.attention_backward(ctx, cfg, P, G, h, y, dlogp)
    ev[2].record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    if it >= 3:
        tf += ev[0].elapsed_time(ev[1]) / N; tb += ev[1].elapsed_time(ev[2]) / N
ctx.profile(True)
for _ in range(5):
    s2s.attention_forward(ctx, cfg, P, h, y)
    s2s.attention_backward(ctx, cfg, P, G, h, y, dlogp)
pr = ctx.profile_read(); ctx.profile(False)
if pr["dec_bwd"][1]:
    print(f"dec_cluster_bwd_kernel: {pr['dec_bwd'][0] / pr['dec_bwd'][1] * 1e3:.1f} us per launch = {pr['dec_bwd'][0] / pr['dec_bwd'][1] * 1e3 / T:.2f} us per decoder step")
if pr["dec_fwd"][1]:
    print(f"dec_cluster_fwd_kernel: {pr['dec_fwd'][0] / pr['dec_fwd'][1] * 1e3:.1f} us per launch = {pr['dec_fwd'][0] / pr['dec_fwd'][1] * 1e3 / T:.2f} us per decoder step (events around the launch)")
else:
    per = sum(pr[k][0] for k in ("attn_fwd", "dense_small")) / 5
    print(f"per-step path: attention + dense launches {per * 1e3:.1f} us per call = {per * 1e3 / T:.2f} us per decoder step (events around each launch)")
print(f"S2S_DEC_CLUSTER={os.environ.get('S2S_DEC_CLUSTER', '1')} B={B} L={L} T={T}: attention_forward {tf:.3f} ms, attention_backward {tb:.3f} ms "
      f"(eager launches); logp checksum {float(logp.double().sum()):.6f} dh checksum {float(dh.double().abs().sum()):.6f}")
