"""GEMM microbenchmark: exact-fp32 SIMT kernel vs the tcgen05 3xTF32 kernel on the projection shapes of cfg2.
Reports the whole call (operand split/transposition pre-passes included) and the GEMM kernel alone.
usage: python benchmarks/gemm_micro.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s

ctx = s2s.Context(0)
TF32_PEAK = 1130.0   # dense TF32 TFLOP/s (B200_PROFILING.md); bf16 measured peak / 2 is the same ballpark
shapes = [("NT", 9600, 1536, 512, "encoder input projection, layers 2-3 (both directions, 3 gates)"),
          ("NT", 9600, 1536, 123, "encoder input projection, layer 1"),
          ("NT", 9600, 512, 512, "Vh = h W_V^T"),
          ("NN", 9600, 512, 1536, "dX = dA W_x"),
          ("NN", 9600, 512, 512, "dh = dVh W_V"),
          ("TN", 1536, 512, 9600, "dW_x += dA^T X"),
          ("TN", 512, 512, 9600, "dW_V += dVh^T h"),
          ("TN", 512, 256, 9600, "dW_h(z,r) += dA^T h_prev"),
          ("NT", 1600, 448, 768, "decoder maxout layer, time-batched"),
          ("NT", 18944, 256, 512, "fit: exactly one round of 148 tiles, 16 slabs"),
          ("NT", 18944, 256, 1024, "fit: one round, 32 slabs"),
          ("NT", 18944, 256, 2048, "fit: one round, 64 slabs"),
          ("NT", 18944, 512, 2048, "fit: two rounds, 64 slabs")]
if os.environ.get("GEMM_MICRO_ONLY"):
    keep = os.environ["GEMM_MICRO_ONLY"].split(",")
    shapes = [s for s in shapes if any(k in s[4] for k in keep)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for form, M, N, K, what in shapes:
    tA, tB = form[0] == "T", form[1] == "T"
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    beta = 1.0 if tA else 0.0
    for impl, name in ((1, "simt fp32"), (2, "tcgen05 3xTF32")):
        if impl == 1 and os.environ.get("GEMM_MICRO_TC_ONLY"):
            continue
        for _ in range(3):
            s2s.gemm(ctx, A, B, tA=tA, tB=tB, beta=beta, C_out=C, impl=impl)
        reps = 10
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record(); s2s.gemm(ctx, A, B, tA=tA, tB=tB, beta=beta, C_out=C, impl=impl); b.record()
        torch.cuda.synchronize()
        call_us = 1e3 * sum(a.elapsed_time(b) for a, b in ev) / reps
        ctx.profile(True)
        for _ in range(5):
            s2s.gemm(ctx, A, B, tA=tA, tB=tB, beta=beta, C_out=C, impl=impl)
        ms, cnt, work = ctx.profile_read()["gemm"]
        ctx.profile(False)
        t = ms / cnt
        tf = 2.0 * M * N * K / t / 1e9
        extra = f" | tensor work 3x = {3 * tf:6.1f} TF/s = {3 * tf / TF32_PEAK * 100:4.1f}% of TF32 peak" if impl == 2 else ""
        print(f"{form} M={M:5d} N={N:5d} K={K:5d} {name:15s} call {call_us:7.1f} us  kernel {t * 1e3:7.1f} us {tf:6.1f} TFLOP/s fp32-equiv{extra}   # {what}", flush=True)
