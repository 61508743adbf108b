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


@pytest.fixture(autouse=True)
def _host_quantile(request, monkeypatch):
    """GPU parity tests compare thresholds with torch.quantile on the HOST: make that comparison independent of whether the host's
    torch contracts the final lerp (tests/helpers.host_quantile_without_contraction; the identity on hosts that do not)."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch

    from tests.helpers import host_quantile_without_contraction
    monkeypatch.setattr(torch, "quantile", host_quantile_without_contraction(torch.quantile))
    yield
