import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on the CPU")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def parity(got: torch.Tensor, ref: torch.Tensor):
    """(max|d| / max|ref|, cosine) -- the two numbers BASELINE.json's tolerance is stated in
    ("max relative error" is normalised by max|ref|: an element-wise ratio is meaningless near zero)."""
    got = got.detach().double().flatten().cpu()
    ref = ref.detach().double().flatten().cpu()
    rel = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
    cos = (torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-30)).item()
    return rel, cos


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of the C-ABI library; builds it when nvcc is present and the .so is stale/missing."""
    from stabletriton_b200 import build as B

    try:
        return B.build(selftest=False)
    except Exception:
        if os.path.exists(B.LIB_PATH):
            return B.LIB_PATH
        raise
