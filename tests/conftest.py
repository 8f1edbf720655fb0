import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/*.cpp with the system g++."""
    from oracle_api import load_oracle
    return load_oracle()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.skip("no CUDA device")
    return 0


@pytest.fixture(scope="session")
def gen_dir(tmp_path_factory):
    """Scratch directory for generated scenes (scenes/gen_scenes.py)."""
    return str(tmp_path_factory.mktemp("scenes"))
