"""Time the generator FRONT END alone — mapping networks + cross-attention stack + adjust_style, i.e.
`Generator.forward(..., return_only_style_latent=True)` (`model_spatial_query.py:626-686`) — forward and
forward+backward, each replayed from a CUDA graph.  It is a serial chain that sits in front of the first
convolution (and behind the last one in the backward pass), so its time adds to the step 1:1.

    python tools/frontend_bench.py [--batch 16] [--json out.json]
Profiling aid, not a bench."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import model_spatial_query as M  # noqa: E402
from transeditor_b200 import lib  # noqa: E402
from transeditor_b200 import model as te_model  # noqa: E402


def timed_graph(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    l0 = lib.launch_count
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    launches = lib.launch_count - l0
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = "cuda"
    rows = []
    for precision in ("bf16", "fp32"):
        te_model.set_precision(precision)
        torch.backends.cuda.matmul.allow_tf32 = precision == "bf16"
        torch.manual_seed(0)
        g = M.Generator(256, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(dev)
        z, p = torch.randn(a.batch, 512, 16, device=dev), torch.randn(a.batch, 512, 16, device=dev)
        params = [q for q in g.parameters()]

        def fwd():
            with torch.no_grad():
                return g(z, p, return_only_style_latent=True)

        def fwdbwd():
            lat = g(z, p, return_only_style_latent=True)
            torch.autograd.grad(lat.square().sum(), [q for q in params if q.requires_grad], allow_unused=True)

        ms_f, n_f = timed_graph(fwd)
        ms_fb, n_fb = timed_graph(fwdbwd)
        rows.append({"precision": precision, "batch": a.batch, "fwd_ms": round(ms_f, 4), "fwd_te_launches": n_f,
                     "fwdbwd_ms": round(ms_fb, 4), "fwdbwd_te_launches": n_fb})
        print(rows[-1], flush=True)
    te_model.set_precision("fp32")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
