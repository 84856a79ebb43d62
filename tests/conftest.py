import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def nb():
    """The product package with its CUDA library loaded (GPU tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ntedit_b200
    ntedit_b200.lib.load()
    return ntedit_b200
