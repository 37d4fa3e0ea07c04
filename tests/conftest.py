import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", params=["ffma", "tc", "tc16"])
def ctx(request):
    """One CUDA context per engine family for the whole GPU session: every test that takes `ctx` runs once on the CUDA-core
    sweeps (exact-FP32 FFMA for SURF, XOR + POPC for ORB), once on the fp32-accumulator tcgen05 sweeps (3xTF32 for SURF, FP8 dot
    product with packed keys for ORB) and once on the 16-bit tcgen05 sweeps (FP16 split for SURF, FP16 accumulators for ORB, slice
    keys resolved by finalize); fails loudly if the library or the device is missing."""
    import easysfm_b200 as esfm
    c = esfm.Context(0)
    ham = {"ffma": "popc", "tc": "tc", "tc16": "tc16"}[request.param]
    c.set_l2_engine(request.param)
    c.set_hamming_engine(ham)
    assert c.l2_engine() == request.param and c.hamming_engine() == ham
    yield c
    c.close()
