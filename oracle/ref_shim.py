"""Import the UNMODIFIED reference model code on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference exists (this container, never the GPU box): it is
used by oracle/make_golden.py to generate tests/golden/*.npz and by the
`-m "not gpu"` tests that pin oracle/te_oracle.py against the reference itself.

The reference is CUDA-only (model_spatial_query.py:630,642 hard-code `.cuda()`, and
utils/op JIT-builds two CUDA extensions at import).  To run its own arithmetic on
CPU we (1) pre-register a `utils.op` module made of oracle/ops_cpu.py — the only
restated part — and (2) turn `Tensor.cuda()` into the identity while the reference
forward runs.  Everything else (Generator, Discriminator, ModulatedConv2d,
AttentionBlock ...) is the reference's own code.
"""
import contextlib
import importlib
import os
import sys
import types

import torch

def _default_root():
    """TE_REFERENCE_ROOT, else /root/reference (the build container), else the unmodified copy that
    tools/install_reference.py placed under baseline/_ref (git-ignored; it travels to the GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("TE_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "model_spatial_query.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _default_root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model_spatial_query.py"))


def load_reference_module():
    """Return the reference's model_spatial_query module (imported under a private name)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "_te_ref_model" in sys.modules:
        return sys.modules["_te_ref_model"]
    from oracle import ops_cpu

    saved = {k: sys.modules.get(k) for k in ("utils", "utils.op", "model_spatial_query")}
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "utils")]
    stub = types.ModuleType("utils.op")
    stub.FusedLeakyReLU = ops_cpu.FusedLeakyReLU
    stub.fused_leaky_relu = ops_cpu.fused_leaky_relu
    stub.upfirdn2d = ops_cpu.upfirdn2d
    sys.modules["utils"] = pkg
    sys.modules["utils.op"] = stub
    sys.modules.pop("model_spatial_query", None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("model_spatial_query")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.modules.pop("model_spatial_query", None)
        for k in ("utils", "utils.op"):
            if saved[k] is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = saved[k]
    sys.modules["_te_ref_model"] = mod
    return mod


def load_reference_psp_encoders():
    """Return the reference's pSp/models/encoders/psp_encoders_new.py module (its `from model_spatial_query import
    EqualLinear` resolved to the REFERENCE's model file, not to this repository's drop-in of the same name)."""
    if "_te_ref_psp_encoders" in sys.modules:
        return sys.modules["_te_ref_psp_encoders"]
    ref_model = load_reference_module()
    saved = {k: sys.modules.get(k) for k in ("model_spatial_query",)}
    sys.modules["model_spatial_query"] = ref_model
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("pSp.models.encoders.psp_encoders_new")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k == "pSp" or k.startswith("pSp.")]:
            sys.modules.pop(k)
    sys.modules["_te_ref_psp_encoders"] = mod
    return mod


@contextlib.contextmanager
def cpu_mode():
    """Neutralise the hard-coded `.cuda()` calls while reference code runs on CPU."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
