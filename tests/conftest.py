import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def product():
    from chafa_b200 import capi
    return capi.load_product()


@pytest.fixture(scope="session")
def reference():
    from chafa_b200 import capi
    return capi.load_reference()


@pytest.fixture(scope="session")
def gpu(product):
    if not product.chafa_b200_available():
        pytest.fail("no CUDA device: -m gpu tests need the B200 box (the product has no CPU fallback)")
    return product
