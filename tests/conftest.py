import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    """max|a-b| / max|b| on fp32 copies — the per-tensor relative error of SURVEY.md §8d."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def cosine(a, b):
    import torch

    a = a.detach().float().cpu().flatten()
    b = b.detach().float().cpu().flatten()
    return float(torch.nn.functional.cosine_similarity(a, b, dim=0))
