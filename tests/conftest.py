import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ref
    ref.lib()
    return ref


@pytest.fixture(scope="session")
def gen():
    from tools import streamgen
    streamgen.lib()
    return streamgen


@pytest.fixture(scope="session")
def emu():
    from tests import hostemu
    hostemu.lib()
    return hostemu
