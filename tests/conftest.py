import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gf-pl-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle (test infrastructure)"""
    import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def ctx():
    """one gfpl context on cuda:0 — fails loudly if the CUDA library or the device is missing"""
    from gfpl import api
    return api.Context(0)
