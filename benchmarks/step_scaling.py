"""Where does the step go?  Times the graph-replayed model fwd+bwd at several (L, T) to separate the
encoder (scales with L) from the decoder time loop (scales with T).  usage: python benchmarks/step_scaling.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import s2s_b200 as s2s

CFG = dict(s2s.CHOROWSKI_TIMIT)
ctx = s2s.Context(0)
P = torch.from_numpy(s2s.init_params(CFG, seed=1)).cuda()
G = torch.zeros_like(P)
B = 32


def run(L, T, reps=8):
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((B, L, CFG["D"])).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.integers(0, CFG["V"], (B, T)).astype(np.int32)).cuda()
    nll = torch.zeros(B, device="cuda")
    for _ in range(4):
        s2s.model_fwdbwd(ctx, CFG, P, G, X, y, nll=nll)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s2s.model_fwdbwd(ctx, CFG, P, G, X, y, nll=nll)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = run(300, 50)
print(f"L=300 T=50 : {base:.3f} ms")
t25 = run(300, 25); t100 = run(300, 100)
print(f"L=300 T=25 : {t25:.3f} ms   T=100: {t100:.3f} ms   -> decoder loop {(t100 - t25) / 75 * 1e3:.1f} us per step (fwd+bwd)")
l150 = run(150, 50); l600 = run(600, 50)
print(f"L=150 T=50 : {l150:.3f} ms   L=600: {l600:.3f} ms   -> {(l600 - l150) / 450 * 1e3:.2f} us per frame-step (3 layers fwd+bwd + L-proportional decoder work)")
