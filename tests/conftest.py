import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def has_gpu():
    # Ask the product library itself (ptf_device_count): the CPU-only suite then does not need torch at all, and a
    # missing / broken libptf_b200.so fails the run loudly instead of turning every GPU test into a silent skip.
    import ctypes

    import ptf_b200
    n = ctypes.c_int32(0)
    ptf_b200._capi.load().ptf_device_count(ctypes.byref(n))
    return n.value > 0


def pytest_collection_modifyitems(config, items):
    # A gpu-marked test on a box without a GPU is skipped, never silently passed on a CPU path.
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
