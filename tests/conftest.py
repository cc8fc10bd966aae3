import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.port()  # builds liboracle_port.so on demand
    return o


@pytest.fixture(scope="session")
def acwm():
    import acwm_pkg
    if not os.path.exists(os.path.join(acwm_pkg.PKG_DIR, "libacwm_b200.so")):
        import __graft_entry__ as g
        g.build()
    return acwm_pkg.load()


@pytest.fixture(scope="session")
def have_ref(oracle):
    return oracle.ref_available()
