import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The real reference compiled into oracle/_ref (absent => skip)."""
    from oracle import pyoracle as po

    if not po.ref_available():
        pytest.skip("oracle/_ref/libpixiu_ref.so not built (no /root/reference here)")
    r = po.Ref()
    yield r
    r.close()
