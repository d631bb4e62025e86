import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    g.build_cuda()  # no-op when libb2r.so is up to date (it travels prebuilt to the GPU box)
    return g.load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import portbind
    portbind.lib()
    return portbind
