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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def sq():
    import stochquant_b200 as m
    if not os.path.exists(m.library_path()):
        m.build()
    return m


@pytest.fixture(scope="session")
def gpu_sq(sq):
    """The product library on a machine that must have a GPU: fail loudly, never skip."""
    L = sq.load()
    n = L.sq_device_count()
    assert n > 0, "gpu-marked test on a machine without a CUDA device: " + (L.sq_last_cuda_error() or b"").decode()
    return sq
