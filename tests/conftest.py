import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built on demand; nvcc cross-compiles without a GPU)."""
    from mj_video_b200 import _lib, build
    if not _lib.lib_path().exists():
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def golden_forward():
    import torch
    return torch.load(os.path.join(GOLD, "tiny_forward.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_llama_forward():
    import torch
    return torch.load(os.path.join(GOLD, "tiny_llama_forward.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_frames():
    with open(os.path.join(GOLD, "frames.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_chat():
    with open(os.path.join(GOLD, "prepare_chat_input.json")) as f:
        return json.load(f)


def reference_available():
    return os.path.isdir("/root/reference/scripts/model")
