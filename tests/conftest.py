import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """libpedk.so and the oracle, built in-tree (nvcc cross-compiles without a GPU)."""
    from ped_b200 import build

    build.build_libpedk()
    build.build_synth()
    build.build_oracle()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from ped_b200 import api

    c = api.Context(0)
    yield c
    c.close()
