"""Shared helpers of the parity tests (synthetic, seeded inputs)."""
import numpy as np


def rel_err(a, b, floor=1e-6):
    """max |a-b| over max |b| (the reference's scale; `floor` guards references that are exactly zero)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def make_batch(cfg, B, Lmax, Tmax, seed, ragged=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, Lmax, cfg["D"])).astype(np.float32)
    lengths = rng.integers(max(1, Lmax // 2), Lmax + 1, B).astype(np.int32) if ragged else np.full(B, Lmax, np.int32)
    tlens = rng.integers(2, Tmax + 1, B).astype(np.int32) if ragged else np.full(B, Tmax, np.int32)
    lengths[0] = Lmax; tlens[0] = Tmax
    labels = rng.integers(0, cfg["V"] - 1, (B, Tmax)).astype(np.int32)
    for b in range(B):
        labels[b, tlens[b] - 1] = cfg["V"] - 1   # EOS last
        X[b, lengths[b]:] = 0.0                  # padding is zeros (finite)
    return X, lengths, labels, tlens


def dev(t, dtype=None):
    import torch
    x = torch.from_numpy(np.ascontiguousarray(t))
    if dtype is not None:
        x = x.to(dtype)
    return x.cuda()
