"""Max-abs error of every precision mode against the reference goldens (tests/golden/*.npz) and, at 1024^2,
against the CPU oracle.  Writes a small table; run on the GPU box:

    python tools/parity_report.py [--out gpurun_out/parity_modes.txt] [--no-1024]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import te_oracle as O  # noqa: E402  (checker only)
import model_spatial_query as M  # noqa: E402
from transeditor_b200 import model as te_model, tc  # noqa: E402

DEV = "cuda"


def models(size, cm):
    t = 2 * int(np.log2(size)) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1)
    d = M.Discriminator(size, channel_multiplier=cm)
    sdg = O.synthetic_state(O.generator_shapes(size, cm))
    g.load_state_dict(sdg, strict=True)
    d.load_state_dict(O.synthetic_state(O.discriminator_shapes(size, cm)), strict=True)
    return g.to(DEV), d.to(DEV), sdg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-1024", action="store_true")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    modes = [("fp32_simt", 0), ("fp32", 2), ("fp32", 3), ("bf16", 0)]
    rows = []
    cases = []
    for name in ("gd32_b4", "gd64_b2", "gd256_b1"):
        with np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) as z:
            gold = {k: z[k] for k in z.files}
        g, d, _ = models(int(gold["size"]), int(gold["cm"]))
        cases.append((name, g, d, torch.from_numpy(gold["z"]).to(DEV), torch.from_numpy(gold["p"]).to(DEV),
                      torch.from_numpy(gold["img"]), torch.from_numpy(gold["d_real"]), torch.from_numpy(gold["real"]).to(DEV)))
    if not a.no_1024:
        g, _, sdg = models(1024, 2)
        gen = torch.Generator().manual_seed(5)
        z, p = torch.randn(1, 512, 16, generator=gen), torch.randn(1, 512, 16, generator=gen)
        with torch.no_grad():
            ref, _ = O.generator_forward(sdg, z, p, 1024)
        cases.append(("g1024_b1 (vs CPU oracle)", g, None, z.to(DEV), p.to(DEV), ref, None, None))
    for mode, planes in modes:
        te_model.set_precision(mode)
        if planes:
            tc.set_split_planes(planes)
        for name, g, d, z, p, img_ref, dreal_ref, real in cases:
            with torch.no_grad():
                img, _, _ = g(z, p)
                e_img = (img.float().cpu() - img_ref).abs().max().item()
                e_d = float("nan")
                if d is not None:
                    e_d = (d(real).float().cpu() - dreal_ref).abs().max().item()
            rows.append("%-10s planes=%d  %-26s img max-abs %.3e (|img| max %.1f)   D(real) logit max-abs %.3e"
                        % (mode, planes, name, e_img, img_ref.abs().max().item(), e_d))
            print(rows[-1], flush=True)
    te_model.set_precision("fp32")
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
