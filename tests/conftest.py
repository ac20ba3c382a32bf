import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
# Loop-back worlds run several ranks' spinning kernels on ONE GPU from many streams.  With the default 8
# hardware work queues two of those streams can share a queue, which serialises a kernel behind the very
# kernel that waits for it.  (One process per GPU — the product configuration — uses 2 streams.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test may assume the C-ABI library and the oracle's C restatement are built."""
    import b200ddp  # noqa: F401
    from b200ddp._build import build_library

    build_library()
    from oracle.build import build

    build()


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
