import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyref

    return pyref.load_oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference REF backend (oracle/_ref); skipped if not built."""
    from oracle import pyref

    if not pyref.have_reference():
        pytest.skip("oracle/_ref/libgrid_ref.so not built (needs /root/reference)")
    return pyref


@pytest.fixture(scope="session")
def b200():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cp2k_b200 import load_b200

    return load_b200()
