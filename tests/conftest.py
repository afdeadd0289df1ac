import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import oracle
    return oracle.load("port")


@pytest.fixture(scope="session")
def ref_oracle():
    """The reference's own sources (oracle/_ref). Built here from /root/reference; on the GPU box only
    the prebuilt library travels."""
    from oracle import oracle
    if not oracle.have("ref") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle.load("ref")
