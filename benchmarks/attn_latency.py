"""Latency anatomy of the forward attention step at the L2-resident benchmark shape (B=32, L=300, S=A=512).
usage: S2S_ATT_DBG=<mask> python benchmarks/attn_latency.py     mask: 1 = no combine, 2 = no Vh stream, 4 = no h tile / context"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s2s_b200 as s2s
B, L, S, A = 32, 300, 512, 512
ctx = s2s.Context(0)
Vh = torch.randn(B, L, S, device="cuda"); h = torch.randn(B, L, A, device="cuda")
q = torch.randn(B, S, device="cuda"); w = torch.randn(S, device="cuda") / S ** 0.5
alpha = torch.empty(B, L, device="cuda"); c = torch.empty(B, A, device="cuda")
for _ in range(5):
    s2s.attn_step_forward(ctx, Vh, h, q, w, alpha=alpha, c=c)
ctx.profile(True)
for _ in range(50):
    s2s.attn_step_forward(ctx, Vh, h, q, w, alpha=alpha, c=c)
ms, cnt, _ = ctx.profile_read()["attn_fwd"]
print(f"S2S_ATT_DBG={os.environ.get('S2S_ATT_DBG', '0')}: attn_fwd {ms / cnt * 1e3:.2f} us (L2-warm, events around the launch)")
