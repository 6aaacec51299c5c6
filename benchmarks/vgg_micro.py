"""Microbenchmark of the VGG front-end (librispeech/model_vgg.lua:23-54) at the LibriSpeech shape of BASELINE configs[3]:
X [B, 3, 1600, 40] -> annotations [B, 796, 512], forward + backward.
usage: python benchmarks/vgg_micro.py [B T F reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s

B, T, F, reps = (int(x) for x in (sys.argv[1:5] + ["8", "1600", "40", "3"][len(sys.argv) - 1:]))
cfg = s2s.VGG_LIBRISPEECH
ctx = s2s.Context(0)
torch.manual_seed(0)
n = s2s.vgg_param_count(cfg, F)
P = (torch.rand(n, device="cuda") * 2 - 1) * 0.03
X = torch.randn(B, 3, T, F, device="cuda")
L = (T - 8) // 2
dh = torch.randn(B, L, cfg["OUT"], device="cuda")
dP = torch.zeros_like(P)
for _ in range(2):
    h = s2s.vgg_forward(ctx, cfg, P, X)
    s2s.vgg_backward(ctx, cfg, P, X, dh, dP=dP)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(reps):
    ev[0].record(); h = s2s.vgg_forward(ctx, cfg, P, X); ev[1].record()
    s2s.vgg_backward(ctx, cfg, P, X, dh, dP=dP); ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]) / reps; tb += ev[1].elapsed_time(ev[2]) / reps
# FLOPs (multiply-add = 2): the four convolutions and the four 1x1 layers, forward; backward = 2x (dX skipped for conv1)
H1, W1, H2, W2 = T - 2, F - 2, T - 4, F - 4
Wp1 = W2 // 2; H3, W3, H4, W4 = H2 - 2, Wp1 - 2, H2 - 4, Wp1 - 4
Wq = W4 // 2; view = cfg["C2"] * Wq
conv = 2 * 9 * (H1 * W1 * 3 * cfg["C1"] + H2 * W2 * cfg["C1"] ** 2 + H3 * W3 * cfg["C1"] * cfg["C2"] + H4 * W4 * cfg["C2"] ** 2)
lin = 2 * L * (view * cfg["HID"] + 2 * cfg["HID"] ** 2 + cfg["HID"] * cfg["OUT"])
fl = B * (conv + lin)
print(f"VGG front-end B={B} T={T} F={F} -> L={L}: forward {tf:.2f} ms ({fl / tf / 1e9:.1f} TFLOP/s fp32-equivalent), "
      f"backward {tb:.2f} ms ({2 * fl / tb / 1e9:.1f} TFLOP/s); {B * T / ((tf + tb) * 1e-3) / 1e3:.1f} K input frames/s fwd+bwd; "
      f"{conv / 1e9:.1f} + {lin / 1e9:.1f} GFLOP per utterance forward")
ctx.profile(True)
h = s2s.vgg_forward(ctx, cfg, P, X)
s2s.vgg_backward(ctx, cfg, P, X, dh, dP=dP)
ms, cnt, work = ctx.profile_read()["gemm"]
ctx.profile(False)
print(f"  GEMM kernels: {ms:.2f} ms over {cnt} launches ({work / ms / 1e9:.1f} TFLOP/s fp32-equivalent); the rest is unfold/fold/ReLU/pooling/operand preparation")
