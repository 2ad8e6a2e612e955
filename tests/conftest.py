import os
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_api():
    import oracle_api as O
    O.build("port")
    if os.path.exists("/root/reference/ad.cl"):
        O.build("ref")
    return O


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context per test session; fails loudly when the extension is missing."""
    from ad4cl_b200 import host
    c = host.Context(0)
    yield c
    c.close()
