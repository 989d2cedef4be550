"""BASELINE configs[4]: pSp encoder + generator inversion forward, 256^2, batch 32, one B200.

Times (CUDA events, inputs resident on the device, random-init weights of the real architecture):
  * `reference-structure`: the encoder module exactly as the reference wires it (30 separate heads, unfolded batch
    norms, f32 / TF32 off) + generator in the fp32 parity mode, eager launches;
  * `fused`: FusedEncoder (folded batch norms, batched heads, bf16 channels-last) + bf16 tcgen05 generator, the whole
    forward replayed from one CUDA graph (InversionPipeline).

    python tools/inversion_bench.py [--batch 32] [--json out.json]
Profiling aid, not the headline bench."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import model_spatial_query as M  # noqa: E402
from transeditor_b200 import model as te_model  # noqa: E402
from transeditor_b200.inversion import GradualStyleEncoder, InversionPipeline  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    enc = GradualStyleEncoder(50, "ir_se").to(dev).eval()
    g = M.Generator(256, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(dev).eval()
    x = torch.rand(a.batch, 3, 256, 256, device=dev) * 2 - 1
    rows = []

    te_model.set_precision("fp32")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def reference_structure():
        with torch.no_grad():
            z, p = enc(x)
            return g(z, p, use_spatial_mapping=False, use_style_mapping=False)[0]

    ms = timed(reference_structure, 3)
    rows.append({"mode": "reference-structure f32 eager", "batch": a.batch, "ms": round(ms, 2),
                 "img_per_s": round(a.batch / ms * 1e3, 1)})
    print(rows[-1], flush=True)

    te_model.set_precision("bf16")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    pipe = InversionPipeline(enc, g, resize=False, graph=True)
    ms = timed(lambda: pipe(x), 10)
    enc_only = timed(lambda: pipe.encoder(x), 10)
    rows.append({"mode": "fused bf16 cuda-graph", "batch": a.batch, "ms": round(ms, 2),
                 "img_per_s": round(a.batch / ms * 1e3, 1), "encoder_only_eager_ms": round(enc_only, 2)})
    print(rows[-1], flush=True)
    te_model.set_precision("fp32")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
