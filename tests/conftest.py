import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "staged: opt-in code paths written in a GPU-less session and not yet run on "
                                       "hardware; skipped unless B200_STAGED=1 (see DESIGN.md §9)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if os.environ.get("B200_STAGED", "0") != "1":
        skip_staged = pytest.mark.skip(reason="staged (not yet run on hardware): set B200_STAGED=1 to run")
        for item in items:
            if "staged" in item.keywords:
                item.add_marker(skip_staged)
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library; (re)built with nvcc when sources changed (cross-compiles without a GPU)."""
    from tinygpt_b200 import build
    return build.build()
