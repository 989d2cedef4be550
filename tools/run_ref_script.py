"""Run one of the reference's own scripts UNMODIFIED against this repository's modules.

    python tools/run_ref_script.py [--ref ROOT] [--workdir DIR] train_spatial_query.py DATA --iter 1 --batch 4 --size 64
    python tools/run_ref_script.py test_spatial_query.py --ckpt out/test/checkpoint/000000.pt --size 64 --sample
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_ref_script.py train_spatial_query.py DATA ...

The script's source stays byte-identical (it is executed with runpy from the reference tree: TE_REFERENCE_ROOT,
baseline/_ref, or /root/reference); only the PROCESS ENVIRONMENT is adapted:
  * sys.path = [this repo, reference tree]: `model_spatial_query` and `utils.op` resolve to this repository's
    drop-ins, `utils.sample`, `utils.distributed`, `utils.dataset`, `our_interfaceGAN.*` to the reference's own files
    (`utils` is a namespace package on both sides);
  * packages this image lacks are stubbed: `matplotlib` (imported, never used by the script), `wandb`, and `lmdb` — as an
    in-memory store of synthetic PNG images behind lmdb's API, so the reference's own `MultiResolutionDataset`
    (utils/dataset.py) runs as written, PIL decode included;
  * `torchvision.utils.save_image(range=...)` (renamed `value_range` since torchvision 0.13);
  * `--local_rank` from $LOCAL_RANK when launched by torchrun (the script predates it);
  * under torchrun: torch >= 1.10's DistributedDataParallel(find_unused_parameters=True) hands the module's outputs
    through an autograd sink that returns NEW tensor objects; the script's path regulariser differentiates one output
    (the image) with respect to another (the latents, train_spatial_query.py:96-98), which only works on the original
    tensors, as torch 1.7 (the authors' version) returned them.  The sink is a pass-through when static_graph is off,
    so it is replaced by the identity — torch 1.7's behaviour; the reducer is prepared in forward either way.
"""
import argparse
import io
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference(explicit=None):
    for cand in (explicit, os.environ.get("TE_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "train_spatial_query.py")):
            return os.path.abspath(cand)
    return None


def _stub_modules(n_images, seed):
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    try:
        import lmdb  # noqa: F401
    except ImportError:
        sys.modules["lmdb"] = _fake_lmdb(n_images, seed)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:  # tensorboard missing: a writer that drops everything
        tb = types.ModuleType("torch.utils.tensorboard")

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def add_scalars(self, *a, **k):
                pass

            def close(self):
                pass
        tb.SummaryWriter = SummaryWriter
        sys.modules["torch.utils.tensorboard"] = tb


def _fake_lmdb(n_images, seed):
    """lmdb's read API over synthetic images: keys '<res>-<index>' -> PNG bytes, 'length' -> count."""
    import numpy as np
    from PIL import Image
    mod = types.ModuleType("lmdb")

    class _Txn:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def get(self, key):
            key = key.decode("utf-8")
            if key == "length":
                return str(n_images).encode("utf-8")
            res, idx = key.split("-")
            rng = np.random.default_rng(seed + int(idx))
            # smooth random fields so the images are not pure noise (irrelevant to the arithmetic, friendlier to PNG)
            small = rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)
            img = Image.fromarray(small).resize((int(res), int(res)), Image.BILINEAR)
            buf = io.BytesIO()
            img.save(buf, format="PNG")
            return buf.getvalue()

    class _Env:
        def begin(self, write=False):
            return _Txn()

    mod.open = lambda path, **kw: _Env()
    return mod


def _patch_save_image():
    import torchvision.utils as tvu
    orig = tvu.save_image

    def save_image(tensor, fp, *a, **kw):
        if "range" in kw:
            kw["value_range"] = kw.pop("range")
        return orig(tensor, fp, *a, **kw)
    tvu.save_image = save_image


def _patch_ddp_sink():
    import torch.nn.parallel.distributed as ddp
    if hasattr(ddp, "_DDPSink"):
        ddp._DDPSink.apply = lambda ddp_weakref, *inputs: tuple(inputs)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", default=None, help="reference tree (default: $TE_REFERENCE_ROOT, baseline/_ref, /root/reference)")
    ap.add_argument("--workdir", default=None, help="directory the script runs in (it writes ./out, ./generation)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"],
                    help="ours: this repository's model_spatial_query / utils.op shadow the reference's; reference: the "
                         "reference's own modules and CUDA extensions (JIT-built for sm_100a) — the same-GPU baseline")
    ap.add_argument("--tf32", type=int, default=None, choices=[0, 1],
                    help="set torch.backends.{cudnn,cuda.matmul}.allow_tf32 before the script runs (default: leave)")
    ap.add_argument("--synthetic-images", type=int, default=4096)
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    ref = find_reference(a.ref)
    if ref is None:
        sys.exit("run_ref_script: no reference tree (set TE_REFERENCE_ROOT or run tools/install_reference.py)")
    script = a.script if os.path.isabs(a.script) else os.path.join(ref, a.script)
    if not os.path.isfile(script):
        sys.exit("run_ref_script: %s not found" % script)
    _stub_modules(a.synthetic_images, seed=1234)
    _patch_save_image()
    if a.tf32 is not None:
        import torch
        torch.backends.cudnn.allow_tf32 = bool(a.tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    if a.impl == "reference":
        sys.path[:0] = [ref]
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(ROOT, "baseline", "_ref", "_te_build"))
    else:
        sys.path[:0] = [ROOT, ref]
        from transeditor_b200 import model as te_model
        te_model.set_precision(a.precision)
    argv = [script] + list(a.args)
    if "LOCAL_RANK" in os.environ and not any(x.startswith("--local_rank") for x in argv) \
            and os.path.basename(script).startswith("train"):
        argv.append("--local_rank=%s" % os.environ["LOCAL_RANK"])
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _patch_ddp_sink()
    if a.workdir:
        os.makedirs(a.workdir, exist_ok=True)
        os.chdir(a.workdir)
    sys.argv = argv
    import model_spatial_query
    mine = os.path.abspath(model_spatial_query.__file__).startswith(os.path.join(ROOT, "model_spatial_query"))
    assert mine == (a.impl == "ours"), "wrong model_spatial_query on the path: %s" % model_spatial_query.__file__
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
