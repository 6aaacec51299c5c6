"""BASELINE.json configs[4]: attention-step microbenchmark sweep over encoder length L and batch B
(scoring + softmax + context, forward and backward), reported as algorithmic HBM GB/s against the
measured peak.  L2 (126 MB) is flushed between timed launches by a 256 MiB write, so the numbers are
HBM numbers even when Vh + h would fit in L2.
usage: python benchmarks/attn_sweep.py [--quick] [--out file.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import s2s_b200 as s2s


def run_sweep(kf=0, quick=False, noflush=False, verbose=True, ctx=None):
    """rows of the sweep (one per (B, L)); kernel time from the library's own CUDA events around each launch"""
    S = A = 512
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ctx = ctx or s2s.Context(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    Ls = [300, 1000] if quick else [100, 200, 300, 500, 1000, 2000]
    Bs = [32, 128] if quick else [1, 8, 32, 128, 256]
    rows = []
    for B in Bs:
        for L in Ls:
            Vh = torch.randn(B, L, S, device="cuda"); h = torch.randn(B, L, A, device="cuda")
            q = torch.randn(B, S, device="cuda"); w = torch.randn(S, device="cuda") / S ** 0.5
            dc = torch.randn(B, A, device="cuda")
            alpha = torch.empty(B, L, device="cuda"); c = torch.empty(B, A, device="cuda")
            dq = torch.empty(B, S, device="cuda"); de = torch.empty(B, L, device="cuda")
            if kf:
                uw = torch.randn(kf, S, device="cuda") * 0.1
                aprev = torch.softmax(torch.randn(B, L, device="cuda"), dim=1)
                dap = torch.empty(B, L, device="cuda")
                fwd = lambda: s2s.attn_step_forward_loc(ctx, Vh, h, q, w, uw, aprev, alpha=alpha, c=c)
                bwd = lambda: s2s.attn_step_backward_loc(ctx, Vh, h, q, w, uw, aprev, alpha, dc, dq=dq, de=de, dalpha_prev=dap)
            else:
                fwd = lambda: s2s.attn_step_forward(ctx, Vh, h, q, w, alpha=alpha, c=c)
                bwd = lambda: s2s.attn_step_backward(ctx, Vh, h, q, w, alpha, dc, dq=dq, de=de)
            for _ in range(3):
                fwd()
                bwd()
            reps = 10
            ctx.profile(True)
            for _ in range(reps):
                if not noflush:
                    flush.fill_(1)
                fwd()
                if not noflush:
                    flush.fill_(2)
                bwd()
            prof = ctx.profile_read()
            ctx.profile(False)
            tf, tb = prof["attn_fwd"][0] / reps, prof["attn_bwd"][0] / reps
            bytes_f = 4.0 * B * (L * S + L * A + 2 * L + S + A)
            bytes_b = 4.0 * B * (L * S + L * A + 6 * L + 2 * S + 2 * A)
            r = dict(B=B, L=L, KF=kf, fwd_us=tf * 1e3, bwd_us=tb * 1e3, fwd_gbs=bytes_f / tf / 1e6, bwd_gbs=bytes_b / tb / 1e6,
                     fwd_frac=bytes_f / tf / 1e6 / peak, bwd_frac=bytes_b / tb / 1e6 / peak, mbytes=bytes_f / 1e6)
            rows.append(r)
            if verbose:
                print(f"B={B:4d} L={L:5d} {bytes_f / 1e6:8.1f} MB  fwd {tf * 1e3:8.1f} us {r['fwd_gbs']:7.0f} GB/s ({r['fwd_frac']:.2f})   "
                      f"bwd {tb * 1e3:8.1f} us {r['bwd_gbs']:7.0f} GB/s ({r['bwd_frac']:.2f})", flush=True)
            del Vh, h
    return dict(peak_gbs=peak, S=S, A=A, KF=kf, l2_flush=not noflush, rows=rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--noflush", action="store_true")
    ap.add_argument("--kf", type=int, default=0, help="location filter size (0 = content attention; 10 = the K=16, k=10 configuration folded to UW[10, S])")
    args = ap.parse_args()
    res = run_sweep(args.kf, args.quick, args.noflush)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
