import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(params=["handoff", "local"])
def decision_mode(request, monkeypatch):
    """Runs a GPU test once per decision mode of the fused step (include/seekr2_b200.h, FLAG_DECIDE_*): the library reads
    SEEKR2_B200_DECIDE when an engine is created and falls back to the hand-off where LOCAL is not eligible."""
    monkeypatch.setenv("SEEKR2_B200_DECIDE", request.param)
    return request.param
