import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and return the path of libvgs_b200.so."""
    import __graft_entry__ as g
    g.build(oracle=True, quiet=True)
    from vgs_svgs_segmentation_b200 import capi
    return capi.SO_PATH
