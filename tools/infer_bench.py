"""Generator inference throughput (BASELINE configs[0]/[3] shapes): images/s of Generator.forward under
no_grad, CUDA events, inputs on the device; fp32 parity mode and bf16 tensor-core mode.

    python tools/infer_bench.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import model_spatial_query as M  # noqa: E402
from transeditor_b200 import model as te_model  # noqa: E402

dev = "cuda"


def run(size, batch, precision, iters):
    te_model.set_precision(precision)
    torch.backends.cuda.matmul.allow_tf32 = precision == "bf16"
    torch.manual_seed(0)
    t = 2 * (size.bit_length() - 1) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(dev).eval()
    z, p = torch.randn(batch, 512, 16, device=dev), torch.randn(batch, 512, 16, device=dev)
    with torch.no_grad():
        for _ in range(3):
            g(z, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g(z, p)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    te_model.set_precision("fp32")
    return {"size": size, "batch": batch, "precision": precision, "ms_per_forward": round(ms, 3),
            "img_per_s": round(batch / ms * 1e3, 1)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rows = [run(256, 16, "bf16", 20), run(1024, 8, "bf16", 10), run(256, 1, "bf16", 20),
            run(256, 16, "fp32", 3), run(1024, 8, "fp32", 2)]
    for r in rows:
        print(r, flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)
