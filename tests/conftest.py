import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# Several fsr1_shard ranks inside ONE process (tests/test_gpu_sharding.py) use 3 streams each, and their flag-waiting kernels
# must never share a hardware work queue with the kernel they wait for: ask for more queues than the default 8.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


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
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
