import os
import sys

import pytest

# Parity tests check every forward synchronously (a capacity overflow is retried inside the call); the deferred,
# non-blocking mode of the training path is exercised by the tests that switch it on explicitly.
os.environ.setdefault("SGR_OVERFLOW_CHECK", "sync")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; if someone runs them without a device, skip instead of erroring.
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
