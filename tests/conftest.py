import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def fvqa_lib():
    """The in-tree C-ABI library, initialised on cuda:0 (GPU tests only)."""
    from flipped_vqa_b200 import _lib
    return _lib.lib()
