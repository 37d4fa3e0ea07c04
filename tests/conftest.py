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
    """One CUDA context per engine family for the whole GPU session: every test that takes `ctx` runs once on the CUDA-core
    sweeps (exact-FP32 FFMA for SURF, XOR + POPC for ORB) and once on the tcgen05 sweeps (3xTF32 for SURF, FP8 +-1 dot product
    for ORB); fails loudly if the library or the device is missing."""
    import easysfm_b200 as esfm
    c = esfm.Context(0)
    c.set_l2_engine(request.param)
    c.set_hamming_engine({"ffma": "popc", "tc": "tc"}[request.param])
    assert c.l2_engine() == request.param and c.hamming_engine() == {"ffma": "popc", "tc": "tc"}[request.param]
    yield c
    c.close()
