"""Microbenchmark of the LSTM sequence path (one direction): persistent cluster recurrence vs the per-frame path.
usage: python benchmarks/lstm_micro.py [B L H Din reps]      (S2S_LSTM_CLUSTER=0 forces the per-frame path)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s

B, L, H, Din, reps = (int(x) for x in (sys.argv[1:6] + ["32", "300", "128", "256", "5"][len(sys.argv) - 1:]))
ctx = s2s.Context(0)
torch.manual_seed(0)
n = s2s.lstm_param_count(Din, H, False)
P = (torch.rand(n, device="cuda") * 2 - 1) / H ** 0.5
x = torch.randn(B, L, Din, device="cuda")
dy = torch.randn(B, L, H, device="cuda")
for _ in range(2):
    y, save = s2s.lstm_seq_forward(ctx, P, x, H)
    s2s.lstm_seq_backward(ctx, P, x, y, save, dy, H)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(reps):
    ev[0].record(); y, save = s2s.lstm_seq_forward(ctx, P, x, H); ev[1].record()
    s2s.lstm_seq_backward(ctx, P, x, y, save, dy, H); ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]) / reps; tb += ev[1].elapsed_time(ev[2]) / reps
ctx.profile(True)
for _ in range(reps):
    y, save = s2s.lstm_seq_forward(ctx, P, x, H)
    s2s.lstm_seq_backward(ctx, P, x, y, save, dy, H)
prof = ctx.profile_read()
ctx.profile(False)
for k in ("gru_fwd", "gru_bwd"):       # the persistent recurrence kernels report in the recurrence classes
    ms, cnt, work = prof[k]
    if cnt:
        print(f"  recurrence kernel ({k[4:]}): {ms / cnt * 1e3:.0f} us/launch = {ms / cnt * 1e3 / L:.2f} us/step")
mode = "per-frame launches" if os.environ.get("S2S_LSTM_CLUSTER") == "0" else "persistent cluster kernels"
print(f"LSTM B={B} L={L} H={H} Din={Din} ({mode}): forward {tf * 1e3:.0f} us = {tf * 1e3 / L:.2f} us/step, "
      f"backward {tb * 1e3:.0f} us = {tb * 1e3 / L:.2f} us/step (whole call: projections, recurrence, weight gradients)")
