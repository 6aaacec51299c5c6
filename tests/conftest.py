import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc32():
    from oracle.oracle import Oracle, build
    build()
    return Oracle("f32")


@pytest.fixture(scope="session")
def orc64():
    from oracle.oracle import Oracle, build
    build()
    return Oracle("f64")


@pytest.fixture(scope="session")
def s2s():
    import s2s_b200
    return s2s_b200


@pytest.fixture(scope="session")
def gctx(s2s):
    """One library context on cuda:0 for the whole GPU test session."""
    ctx = s2s.Context(0)
    yield ctx
    ctx.close()


def rel_err(a, b):
    """max |a-b| / max |b| (scale of the reference)"""
    import numpy as np
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
