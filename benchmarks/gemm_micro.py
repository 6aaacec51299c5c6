"""GEMM microbenchmark: exact-fp32 SIMT kernel vs the tcgen05 3xTF32 kernel on the projection shapes.
usage: python benchmarks/gemm_micro.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s

ctx = s2s.Context(0)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    peak = 1590.0
shapes = [(9600, 1536, 512, "encoder input projection, layers 2-3 (both directions, 3 gates)"),
          (9600, 512, 512, "Vh = h W_V^T"),
          (9600, 512, 1536, "dX = dA W_x (after transposing W_x)"),
          (1536, 512, 9600, "dW_x = dA^T X (after transposing both)"),
          (1600, 448, 768, "decoder maxout layer, time-batched")]
for M, N, K, what in shapes:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
    for impl, name in ((1, "simt fp32"), (2, "tcgen05 3xTF32")):
        for _ in range(2):
            s2s.gemm(ctx, A, B, tB=True, C_out=C, impl=impl)
        ctx.profile(True)
        for _ in range(5):
            s2s.gemm(ctx, A, B, tB=True, C_out=C, impl=impl)
        ms, cnt, work = ctx.profile_read()["gemm"]
        ctx.profile(False)
        t = ms / cnt
        tf = 2.0 * M * N * K / t / 1e9
        extra = f"  tensor-pipe work 3x -> {3 * tf:7.1f} TF/s = {3 * tf / peak * 100:4.1f}% of bf16 peak/2 equiv" if impl == 2 else ""
        print(f"M={M:5d} N={N:5d} K={K:5d} {name:15s} {t * 1e3:8.1f} us {tf:7.1f} TFLOP/s (fp32-equivalent){extra}   # {what}")
