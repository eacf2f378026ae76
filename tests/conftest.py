import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return np.load(os.path.join(d, name + ".npz"))
    return load


_poison_round = [0]


@pytest.fixture(autouse=True)
def _poisoned_shared_memory(request):
    """Every GPU test starts from SMs whose shared memory holds a known non-zero pattern (0xFF / 0x7F alternating):
    a kernel that reads shared memory it never wrote then fails in the test that exercises it, not at random after
    whatever ran before (a letterbox fault of exactly that kind was found by the fuzz test)."""
    if "gpu" in request.keywords:
        from face_detection_multi_scale_b200 import ops
        _poison_round[0] += 1
        ops.debug_poison_smem(0xFF if _poison_round[0] & 1 else 0x7F)
    yield
