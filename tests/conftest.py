import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return O.ref()


@pytest.fixture(scope="session")
def checker():
    """Strongest checker available: the real reference if oracle/_ref travelled, else the pinned C port."""
    from oracle import oracle as O
    return O.best()


@pytest.fixture(scope="session")
def capi():
    from dashing_b200 import capi as c
    return c


@pytest.fixture(scope="session")
def gpu(capi):
    if capi.device_count() < 1:
        pytest.fail("`-m gpu` test selected but libdashing_b200 sees no CUDA device")
    return capi
