import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import _libs
    return _libs.oracle(nt=os.cpu_count() or 1)


@pytest.fixture(scope="session")
def ref():
    import _libs
    r = _libs.reference(nt=os.cpu_count() or 1)
    if r is None:
        pytest.skip("compiled reference (oracle/_ref/libecsimd_ref.so) not available on this host")
    return r


@pytest.fixture(scope="session")
def eng():
    """the CUDA engine through its C ABI; fails (does not skip) when the library is missing"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ecsimd_b200
    from ecsimd_b200 import host
    ecsimd_b200.init(0)
    return host
