import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

BONES = ROOT / "tests" / "golden" / "bones"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def bone_obbs():
    from shoulder_b200.meshio import PcaObb
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = PcaObb(BONES / f"{name}.npz")
        return cache[name]
    return get


@pytest.fixture(scope="session")
def gpu_backend():
    from shoulder_b200 import _lib
    _lib.init(int(os.environ.get("LOCAL_RANK", "0")))
    return _lib
