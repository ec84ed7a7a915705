import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def tables():
    from tests import util
    return util.load_tables()


@pytest.fixture(scope="session")
def gpu_ctx():
    import coati_b200
    ctx = coati_b200.Context(0)   # raises if the CUDA library or the device is missing: no fallback
    yield ctx
    ctx.close()
