import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Backend:
    def __init__(self, device):
        import torch
        self.device = torch.device(device)
        self.name = device

    def dev(self, t):
        return t.to(self.device)


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "emulated"])
def be(request):
    """Parity tests run twice: on the real kernels (`-m gpu`, B200 box) and, in the no-GPU container, on the
    test-only CPU emulation of the C ABI (tests/cpu_emulation.py) so host logic and formulas are covered."""
    if request.param == "gpu":
        yield Backend("cuda")
    else:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from cpu_emulation import Emulated
        with Emulated():
            yield Backend("cpu")
