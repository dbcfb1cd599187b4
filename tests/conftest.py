import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def synthetic_sd():
    """Seeded reference-format netG state_dict (982 tensors, 267 M params), shared by the whole session."""
    from ctrlhair_b200 import synth
    return synth.make_state_dict()


@pytest.fixture(scope="session")
def lib():
    from ctrlhair_b200 import _lib
    return _lib.load()
