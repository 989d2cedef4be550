"""Where do the non-te (torch) kernels of one eager training step come from?

Runs one plain step (D step + G step) under a TorchDispatchMode and prints every aten op grouped by
(op, tensor shapes, forward call site / "backward") with call counts and the bytes it touches.  Large ops cost
their bytes, tiny ones a launch each.  Profiling aid, not a bench.

    python tools/glue_profile.py [--heavy] > gpurun_out/glue_profile.txt
"""
import argparse
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode
from torch.utils._pytree import tree_flatten

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SKIP = ("aten.view", "aten._unsafe_view", "aten.reshape", "aten.permute", "aten.transpose", "aten.t.", "aten.detach",
        "aten.slice", "aten.select", "aten.expand", "aten.unsqueeze", "aten.squeeze", "aten.as_strided",
        "aten.alias", "aten.empty", "aten.unbind", "aten.split", "aten._local_scalar_dense", "aten.is_",
        "aten.sym_", "aten.stride", "aten.size", "aten.unfold", "aten.lift_fresh", "aten.new_empty",
        "aten.empty_like", "aten.view_as", "aten.chunk", "aten.narrow", "aten.result_type", "aten.item",
        "aten._reshape_alias", "aten.set_")


class Logger(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.agg = collections.defaultdict(lambda: [0, 0])

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = str(func)
        if name.startswith(SKIP):
            return out
        flat_in, _ = tree_flatten((args, kwargs or {}))
        flat_out, _ = tree_flatten(out)
        tens = [t for t in flat_in + flat_out if isinstance(t, torch.Tensor) and t.is_cuda]
        if not tens:
            return out
        byts = sum(t.numel() * t.element_size() for t in tens)
        shapes = ",".join("%s%s" % (str(t.dtype)[6:8], list(t.shape)) for t in flat_in if isinstance(t, torch.Tensor))
        site = "backward"
        for fr in reversed(traceback.extract_stack(limit=40)):
            if "transeditor_b200/" in fr.filename and not fr.filename.endswith("lib.py"):
                site = "%s:%d %s" % (os.path.basename(fr.filename), fr.lineno, fr.name)
                break
        k = (name, shapes[:110], site)
        self.agg[k][0] += 1
        self.agg[k][1] += byts
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heavy", action="store_true", help="profile iteration 0 (R1 + path-length phases too)")
    ap.add_argument("--top", type=int, default=400)
    a = ap.parse_args()
    from transeditor_b200 import lib
    from transeditor_b200 import model as te_model
    from transeditor_b200.train_step import TrainConfig, Trainer
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    te_model.set_precision("bf16")
    cfg = TrainConfig(size=256, batch=16)
    tr = Trainer(cfg, dev, seed=0)
    real = (torch.rand(cfg.batch, 3, cfg.size, cfg.size) * 2 - 1).to(dev)
    for _ in range(2):
        tr.step(real)
    tr.iteration = 0 if a.heavy else 1
    torch.cuda.synchronize()
    l0 = lib.launch_count
    log = Logger()
    with log:
        tr.step(real)
    torch.cuda.synchronize()
    rows = sorted(log.agg.items(), key=lambda kv: -(kv[1][1] + kv[1][0] * 12e6))  # ~2 us launch == 12 MB of traffic
    n = sum(v[0] for v in log.agg.values())
    b = sum(v[1] for v in log.agg.values())
    print("aten ops %d, bytes touched %.1f MB; te launches %d" % (n, b / 1e6, lib.launch_count - l0))
    print("%6s %10s  %-28s %-46s %s" % ("calls", "MB", "op", "site", "input shapes"))
    for (op, shapes, site), (cnt, byts) in rows[:a.top]:
        print("%6d %10.1f  %-28s %-46s %s" % (cnt, byts / 1e6, op[5:33], site[:46], shapes))
    bysite = collections.defaultdict(lambda: [0, 0])
    for (op, shapes, site), (cnt, byts) in log.agg.items():
        bysite[site][0] += cnt
        bysite[site][1] += byts
    print("\nby site:")
    for site, (cnt, byts) in sorted(bysite.items(), key=lambda kv: -(kv[1][1] + kv[1][0] * 12e6))[:80]:
        print("%6d %10.1f  %s" % (cnt, byts / 1e6, site))


if __name__ == "__main__":
    main()
