import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def port():
    from oracle import load
    return load("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref); None where it was never built (e.g. a fresh GPU box
    without /root/reference and without the prebuilt .so)."""
    from oracle import load
    return load("reference")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz")))


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI.  Fails (does not skip) when the library or a device is missing:
    the -m gpu tests must never pass on a fallback."""
    from crypto_recommendation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
