"""The reference's OWN CUDA ops and nn.Modules on the GPU.  TEST / BASELINE INFRASTRUCTURE ONLY.

Needs an unmodified reference tree (TE_REFERENCE_ROOT, baseline/_ref — installed by tools/install_reference.py and
shipped to the GPU box by gpurun — or /root/reference).  Used by
  * tests/test_gpu_reference_ops.py: te_upfirdn2d / te_fused_bias_act against the reference's kernels
    (utils/op/upfirdn2d_kernel.cu, fused_bias_act_kernel.cu JIT-built for sm_100a exactly as utils/op/fused_act.py:9-15
    and utils/op/upfirdn2d.py:8-14 build them), forward, backward and double backward;
  * tests/test_gpu_reference_model.py: this repository's Generator / Discriminator against the reference's classes on
    the same GPU with the same state_dict;
  * tools/reference_gpu_baseline.py and `bench.py --impl reference`: the reference G+D step (cuDNN) on the same B200.
The extension build directory is baseline/_ref/_te_build (prebuilt in the container by __graft_entry__.build() with
TORCH_CUDA_ARCH_LIST=10.0a — nvcc cross-compiles without a GPU — so the GPU box only loads it).
"""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.isdir("/root/repo") and os.path.samefile(ROOT, "/root/repo"):
    ROOT = "/root/repo"  # the same absolute paths in the container and on the GPU box: the prebuilt ninja files stay valid


def reference_root():
    for cand in (os.environ.get("TE_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "utils", "op", "fused_act.py")):
            return os.path.abspath(cand)
    return None


def build_dir():
    d = os.path.join(ROOT, "baseline", "_ref", "_te_build")
    os.makedirs(d, exist_ok=True)
    return d


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_ops():
    """(fused_act module, upfirdn2d module) of the reference, their CUDA extensions JIT-built / loaded."""
    if "_te_ref_fused_act" in sys.modules:
        return sys.modules["_te_ref_fused_act"], sys.modules["_te_ref_upfirdn2d"]
    ref = reference_root()
    if ref is None:
        raise RuntimeError("no reference tree (TE_REFERENCE_ROOT / baseline/_ref / /root/reference)")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["TORCH_EXTENSIONS_DIR"] = build_dir()
    fa = _load_file("_te_ref_fused_act", os.path.join(ref, "utils", "op", "fused_act.py"))
    up = _load_file("_te_ref_upfirdn2d", os.path.join(ref, "utils", "op", "upfirdn2d.py"))
    return fa, up


def load_reference_model():
    """The reference's model_spatial_query module wired to ITS OWN CUDA ops (imported under a private name)."""
    if "_te_ref_gpu_model" in sys.modules:
        return sys.modules["_te_ref_gpu_model"]
    ref = reference_root()
    fa, up = load_reference_ops()
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.op", "model_spatial_query")}
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(ref, "utils")]
    opm = types.ModuleType("utils.op")
    opm.FusedLeakyReLU, opm.fused_leaky_relu, opm.upfirdn2d = fa.FusedLeakyReLU, fa.fused_leaky_relu, up.upfirdn2d
    sys.modules["utils"], sys.modules["utils.op"] = pkg, opm
    sys.modules.pop("model_spatial_query", None)
    try:
        mod = _load_file("_te_ref_gpu_model", os.path.join(ref, "model_spatial_query.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def prebuild():
    """Compile the reference's two extensions for sm_100a into baseline/_ref/_te_build (no GPU needed)."""
    load_reference_ops()
    return build_dir()


if __name__ == "__main__":
    print(prebuild())
