import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ctx():
    from foldseek_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture()
def sockdir():
    """a short directory for the gpuserver's Unix-domain socket (sun_path holds 108 bytes; pytest's tmp_path can be much longer)"""
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="fsb", dir="/tmp")
    yield d
    shutil.rmtree(d, ignore_errors=True)
