import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.build(ref=os.path.isdir("/root/reference"))
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def alens_lib():
    import alens_b200

    return alens_b200.Library.get()


@pytest.fixture()
def ctx(alens_lib):
    import alens_b200

    c = alens_b200.Context(device=0)
    yield c
    c.close()
