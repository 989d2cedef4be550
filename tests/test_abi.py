"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares;
the product path fails loudly (no CPU fallback).  CPU only — no kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from tests.conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "te_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(te_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from transeditor_b200 import lib
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(handle, n), "missing export " + n
    assert set(lib._SIGNATURES) == set(names), "ctypes table and header disagree"
    assert lib.load().te_version() >= 1000


def test_no_cpu_fallback():
    import model_spatial_query as M
    from utils.op import fused_leaky_relu, upfirdn2d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fused_leaky_relu(torch.zeros(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    d = M.Discriminator(32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.zeros(4, 3, 32, 32))


def test_missing_library_is_loud(monkeypatch):
    from transeditor_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libte_b200.so")
    with pytest.raises(RuntimeError, match="is missing"):
        lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "transeditor_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("(the oracle", ""), fn + " references the oracle"
    for fn in ("model_spatial_query.py", "utils/op/__init__.py", "utils/op/fused_act.py", "utils/op/upfirdn2d.py"):
        assert "oracle" not in open(os.path.join(ROOT, fn)).read()
