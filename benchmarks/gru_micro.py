"""Microbenchmark of the persistent cluster GRU sequence kernels (one encoder layer, both directions).
usage: python benchmarks/gru_micro.py [B L H Din reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s

B, L, H, Din, reps = (int(x) for x in (sys.argv[1:6] + ["32", "300", "256", "512", "5"][len(sys.argv) - 1:]))
ctx = s2s.Context(0)
torch.manual_seed(0)
W = (torch.rand(6, H, H + Din, device="cuda") * 2 - 1) / (H + Din) ** 0.5
x = torch.randn(B, L, Din, device="cuda")
dy = torch.randn(B, L, 2 * H, device="cuda")
lengths = torch.full((B,), L, dtype=torch.int32, device="cuda") if os.environ.get("GRU_MICRO_LENGTHS") else None
for _ in range(3):                       # warm-up: clocks, kernel attributes, workspaces
    y, save = s2s.gru_seq_forward(ctx, W, x, ndir=2, lengths=lengths)
    dx, dW = s2s.gru_seq_backward(ctx, W, x, y, save, dy, ndir=2, lengths=lengths)
ctx.profile(True)
for _ in range(reps):
    y, save = s2s.gru_seq_forward(ctx, W, x, ndir=2, lengths=lengths)
    dx, dW = s2s.gru_seq_backward(ctx, W, x, y, save, dy, ndir=2, lengths=lengths)
prof = ctx.profile_read()
for k in ("gru_fwd", "gru_bwd", "gemm"):
    ms, cnt, work = prof[k]
    if cnt:
        print(f"{k}: {ms / cnt * 1e3:.1f} us/launch, {ms / cnt * 1e3 / L:.2f} us/step, {cnt} launches")
