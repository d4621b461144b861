import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import imagesequenceregistrationfor6dposeestimationlabeling_b200 as p
    return p


@pytest.fixture(scope="session")
def gpu(pkg):
    """The package with the CUDA library loaded on cuda:0; fails loudly if either is absent."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    torch.cuda.set_device(0)
    pkg._lib.load()
    return pkg
