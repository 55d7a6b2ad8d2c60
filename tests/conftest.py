import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def oracle():
    from tests import support
    return support.Oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import support
    r = support.Reference.try_load()
    if r is None:
        pytest.skip("oracle/_ref/libramref.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def emul():
    from tests import support
    return support.Emulator()


@pytest.fixture(scope="session")
def ramcu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from repeatafterme_b200 import RamCu
    return RamCu(0)
