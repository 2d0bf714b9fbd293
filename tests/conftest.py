import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lightmetrica-v3_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind oracle/_ref (built by oracle/Makefile where /root/reference exists)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    ref.init(0)
    return ref
