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
def oracle():
    from common import oracle_lib
    return oracle_lib()


@pytest.fixture(scope="session")
def product():
    """The CUDA library through the C ABI.  Fails (does not skip) when it is missing:
    a GPU test that silently ran on something else would void the parity claim."""
    from geometricvofext_b200 import capi
    return capi.load_product()
