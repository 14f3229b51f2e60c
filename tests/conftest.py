"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs on the CPU-only build container; `-m gpu` runs on a B200 box, where
/root/reference does not exist (nothing here reads it).
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")
for p in (REPO, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(REPO, "tests", "golden", "activation1d_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_cases(golden):
    names = sorted({k.split("/")[0] for k in golden if "/" in k})
    out = {}
    for n in names:
        out[n] = {k.split("/", 1)[1]: v for k, v in golden.items() if k.startswith(n + "/")}
    return out
