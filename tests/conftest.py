import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", params=["ffma", "tc"])
def ctx(request):
    """One CUDA context per L2 engine for the whole GPU session (every test that takes `ctx` runs once on the
    exact-FP32 FFMA sweep and once on the tcgen05 3xTF32 sweep; the Hamming path is the same kernel in both);
    fails loudly if the library or the device is missing."""
    import easysfm_b200 as esfm
    c = esfm.Context(0)
    c.set_l2_engine(request.param)
    assert c.l2_engine() == request.param
    yield c
    c.close()
