import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def drt():
    import drt_b200
    return drt_b200


@pytest.fixture(scope="session")
def ctx(drt):
    """One CUDA context for the whole GPU session.  No fallback: if the library
    is missing or there is no device this raises, it does not skip."""
    c = drt.Context(0)
    yield c
    c.close()
