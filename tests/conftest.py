import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test (building the checker's target is not using it)."""
    from transeditor_b200 import build
    build.build()


@pytest.fixture(autouse=True)
def _fp32_math():
    """The oracle runs with TF32 disabled; every test starts in the default precision mode."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from transeditor_b200 import model as te_model
    te_model.set_precision("fp32")
    yield
    te_model.set_precision("fp32")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def small(t):
    """Same sampling as oracle/make_golden.py::_small."""
    a = t.detach().cpu().numpy().reshape(-1)
    return a[::97][:8192].copy() if a.size > 65536 else a.copy().reshape(tuple(t.shape))


@pytest.fixture
def cpu_emulation(monkeypatch):
    """Route transeditor_b200.lib's kernel calls to tests/emu.py so the HOST logic (autograd
    wiring, geometry algebra, module orchestration) can be checked without a GPU.  Test-only:
    the product has no such path."""
    from tests import emu
    emu.install(monkeypatch)
    yield
