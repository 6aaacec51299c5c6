"""MN-major tcgen05 operands (S2S_TC_MN=1): relative error of the NN / TN forms against float64, with a few values on failure."""
import numpy as np
import torch

import s2s_b200 as s2s
from tests.util import dev, rel_err

ctx = s2s.Context(0)
rng = np.random.default_rng(0)
for (tA, tB, M, N, K) in [(False, False, 256, 256, 64), (False, False, 9600, 512, 768), (True, False, 256, 256, 64), (True, False, 1536, 512, 9600),
                          (True, False, 768, 124, 9600), (True, False, 100, 36, 1600)]:
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if not tB else (N, K)).astype(np.float32)
    ref = (A.T if tA else A).astype(np.float64) @ (B if not tB else B.T).astype(np.float64)
    Cd = torch.full((M, N), 7.0, device="cuda")
    s2s.gemm(ctx, dev(A), dev(B), tA=tA, tB=tB, C_out=Cd, impl=2)
    torch.cuda.synchronize()
    C = Cd.cpu().numpy()
    e = rel_err(C, ref)
    print(("TN" if tA else "NN"), M, N, K, "rel err %.3e" % e, "" if e < 2e-5 else "\n  got %s\n  ref %s" % (C[0, :6], ref[0, :6]))
ctx.close()
