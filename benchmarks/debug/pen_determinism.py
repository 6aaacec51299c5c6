"""Is the decoder forward run-to-run deterministic, and how close to zero is the monotonicity-penalty gate?  (diagnostic for
tests/test_gpu_timed_path.py::test_cluster_decoder_backward_equals_per_step_chain[K=16, lambda=0.03])"""
import numpy as np
import torch

import s2s_b200 as s2s
from oracle.oracle import Oracle, build, init_params
from tests.util import dev

CFG2 = dict(D=123, H=256, NL=0, S=512, ST=256, V=62, K=16, KF=10, M=64, MW=7)
B, L, T = 32, 300, 50
build()
orc = Oracle("f64")
P = dev(init_params(CFG2, seed=1234, dtype=np.float64, oracle=orc), torch.float32)
rng = np.random.default_rng(7)
h = dev(rng.standard_normal((B, L, 512)) * 0.5, torch.float32)
y = dev(rng.integers(0, CFG2["V"] - 1, (B, T)).astype(np.int32))
outs = []
for run in range(6):
    ctx = s2s.Context(0)
    s2s.attention_forward(ctx, CFG2, P, h, y, lam=0.03)
    a = s2s.attention_get(ctx, 0, (B, T, L)).cpu().numpy().copy()
    p = s2s.attention_get(ctx, 3, (B, T)).cpu().numpy().copy()
    outs.append((a, p))
    ctx.close()
a0, p0 = outs[0]
print("penalty: min |p| = %.3e, values with |p| < 1e-4: %d of %d" % (np.abs(p0).min(), (np.abs(p0) < 1e-4).sum(), p0.size))
for i, (a, p) in enumerate(outs[1:], 1):
    print("run %d: alpha bit-identical %s (max diff %.3e), penalty bit-identical %s, gate flips %d" %
          (i, np.array_equal(a, a0), np.abs(a - a0).max(), np.array_equal(p, p0), ((p > 0) != (p0 > 0)).sum()))
