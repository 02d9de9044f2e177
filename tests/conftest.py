import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import gtb  # noqa: E402,F401  (registers tinyllama_cpp_b200)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test (full-size reference runs)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.ref_available() and not Path("/root/reference/tinyllama.cpp").exists():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle.ref()


@pytest.fixture(scope="session")
def checker():
    """The strongest CPU checker present: the real reference if its prebuilt library travelled, else the port."""
    import oracle
    return oracle.best()
