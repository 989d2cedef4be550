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


def run_graphed(size, batch, precision, iters, mapped=False):
    """Same forward replayed from a CUDA graph (transeditor_b200.inference.GraphedGenerator); `mapped` bypasses the
    two mapping networks like the editing scripts do."""
    from transeditor_b200.inference import GraphedGenerator, edit_frames
    te_model.set_precision(precision)
    torch.backends.cuda.matmul.allow_tf32 = precision == "bf16"
    torch.manual_seed(0)
    t = 2 * (size.bit_length() - 1) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(dev).eval()
    z, p = torch.randn(batch, 512, 16, device=dev), torch.randn(batch, 512, 16, device=dev)
    flags = dict(use_style_mapping=False, use_spatial_mapping=False) if mapped else {}
    gg = GraphedGenerator(g, batch, **flags)
    gg(z, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gg(z, p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    row = {"size": size, "batch": batch, "precision": precision, "mode": "cuda-graph" + ("+mapped-codes" if mapped else ""),
           "ms_per_forward": round(ms, 3), "img_per_s": round(batch / ms * 1e3, 1)}
    if mapped and batch == 1:
        bz = torch.nn.functional.normalize(torch.randn(1, 512, device=dev), dim=1)
        zt, pt = z.transpose(1, 2).contiguous(), p.transpose(1, 2).contiguous()
        edit_frames(gg, zt, pt, z_boundary=bz, p_boundary=bz, z_distance=3.0, p_distance=3.0, steps=10)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        frames = edit_frames(gg, zt, pt, z_boundary=bz, p_boundary=bz, z_distance=3.0, p_distance=3.0, steps=10)
        host = frames.cpu()  # the frames a script would hand to PIL
        row["edit_10_frames_to_host_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        assert host.shape[0] == 10
    te_model.set_precision("fp32")
    return row


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
    return {"size": size, "batch": batch, "precision": precision, "mode": "eager", "ms_per_forward": round(ms, 3),
            "img_per_s": round(batch / ms * 1e3, 1)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rows = [run(256, 16, "bf16", 20), run(1024, 8, "bf16", 10), run(256, 1, "bf16", 20),
            run_graphed(256, 16, "bf16", 20), run_graphed(1024, 8, "bf16", 10), run_graphed(256, 1, "bf16", 50),
            run_graphed(256, 1, "bf16", 50, mapped=True), run_graphed(1024, 1, "bf16", 20, mapped=True),
            run(256, 16, "fp32", 3), run(1024, 8, "fp32", 2)]
    for r in rows:
        print(r, flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)
