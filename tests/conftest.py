import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; if one is selected anyway without a device, fail loudly
    # (no silent CPU fallback exists in the product path).
    pass


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o
