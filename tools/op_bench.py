"""Per-kernel roofline numbers at BASELINE's flagship shapes (CUDA events, L2 flushed between launches).

    python tools/op_bench.py [--json gpurun_out/op_bench.json] [--only NAME]
Also the target command for `ncu --set full -k regex:...` captures (keep --iters small there)."""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import op  # noqa: E402
from utils.op import fused_leaky_relu, upfirdn2d  # noqa: E402

dev = "cuda"
PEAKS = {"hbm_gbs": 6577.7, "tf": 1723.6}
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    with open(pk) as f:
        d = json.load(f)
    PEAKS = {"hbm_gbs": d["hbm_gbs"], "tf": d["bf16_tflops"]}


def timeit(fn, iters):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    res = {}
    fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.]))
    fir = (fir / fir.sum()).to(dev)

    def want(n):
        return a.only is None or a.only in n

    if True:
        if want("upfirdn2d_blur_f32"):
            x = torch.randn(16, 128, 257, 257, device=dev)
            ms = timeit(lambda: upfirdn2d(x, fir, pad=(1, 1)), a.iters)
            byts = 4 * 16 * 128 * (257 * 257 + 256 * 256)
            res["upfirdn2d_blur_f32 [16,128,257,257]->256^2"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                  "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x
        if want("upfirdn2d_blur_dpad2_f32"):
            x = torch.randn(16, 128, 256, 256, device=dev)
            ms = timeit(lambda: upfirdn2d(x, fir, pad=(2, 2)), a.iters)
            byts = 4 * 16 * 128 * (257 * 257 + 256 * 256)
            res["upfirdn2d_blur_f32 [16,128,256,256]->257^2 (D)"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                     "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x
        if want("upfirdn2d_blur_bf16_nhwc"):
            x = torch.randn(16, 128, 257, 257, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: upfirdn2d(x, fir, pad=(1, 1)), a.iters)
            byts = 2 * 16 * 128 * (257 * 257 + 256 * 256)
            res["upfirdn2d_blur_bf16_nhwc [16,257,257,128]->256^2"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                       "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x
        if want("upfirdn2d_blur_bf16_nhwc"):
            x = torch.randn(16, 128, 256, 256, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: upfirdn2d(x, fir, pad=(2, 2)), a.iters)
            byts = 2 * 16 * 128 * (257 * 257 + 256 * 256)
            res["upfirdn2d_blur_bf16_nhwc [16,256,256,128]->257^2 (D)"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                           "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            x = torch.randn(16, 512, 65, 65, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: upfirdn2d(x, fir, pad=(1, 1)), a.iters)
            byts = 2 * 16 * 512 * (65 * 65 + 64 * 64)
            res["upfirdn2d_blur_bf16_nhwc [16,65,65,512]->64^2"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                    "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x
        if want("scale_bc"):
            x = torch.randn(16, 128, 256, 256, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            sc = torch.rand(16, 128, device=dev) + 0.5
            ms = timeit(lambda: op.scale_bc(x, sc), a.iters)
            byts = 2 * 2 * x.numel()
            res["scale_bc_bf16 [16,256,256,128]"] = {"ms": ms, "GB/s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            from transeditor_b200.op import DotBC
            ms = timeit(lambda: DotBC.apply(x, x), a.iters)
            res["dot_bc_bf16 [16,256,256,128]"] = {"ms": ms, "GB/s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x
        if want("fused_bias_act_f32"):
            x = torch.randn(16, 128, 256, 256, device=dev)
            b = torch.randn(128, device=dev)
            ms = timeit(lambda: fused_leaky_relu(x, b), a.iters)
            byts = 2 * 4 * x.numel()
            res["fused_bias_act_f32 [16,128,256,256]"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                          "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: fused_leaky_relu(xb, b.to(torch.bfloat16)), a.iters)
            byts = 2 * 2 * x.numel()
            res["fused_bias_act_bf16_nhwc [16,128,256,256]"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            gy = torch.randn_like(xb)
            xr = xb.clone().requires_grad_(True)
            br = b.to(torch.bfloat16).clone().requires_grad_(True)
            yb = fused_leaky_relu(xr, br)
            ms = timeit(lambda: torch.autograd.grad(yb, (xr, br), gy, retain_graph=True), a.iters)
            byts = 3 * 2 * x.numel()
            res["fused_bias_act_bwd_bf16_nhwc [16,128,256,256]"] = {"ms": ms, "GB/s": byts / ms / 1e6,
                                                                    "frac_hbm": byts / ms / 1e6 / PEAKS["hbm_gbs"]}
            del x, xb, gy, xr, yb
        if want("conv_simt"):
            x = torch.randn(16, 128, 256, 256, device=dev)
            w = torch.randn(128, 128, 3, 3, device=dev) / 34
            ms = timeit(lambda: op.conv2d_fused(x, w, padding=1), max(3, a.iters // 4))
            fl = 2.0 * 16 * 256 * 256 * 128 * 128 * 9
            res["conv2d_simt_f32 128->128@256^2 B16"] = {"ms": ms, "TFLOP/s": fl / ms / 1e9}
            del x
        for (b, h, cin, cout) in ((16, 256, 128, 128), (16, 128, 256, 256), (16, 64, 512, 512)):
            name = f"conv2d_tc_bf16 {cin}->{cout}@{h}^2 B{b}"
            if not want("conv_tc"):
                continue
            x = torch.randn(b, cin, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            wp = op.pack_weight_tc(torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9))
            bias = torch.randn(cout, device=dev)
            osc = torch.rand(b, cout, device=dev) + 0.5
            ms = timeit(lambda: op.conv2d_tc(x, wp, 3, osc, bias, True), a.iters)
            fl = 2.0 * b * h * h * cin * cout * 9
            byts = 2.0 * (b * h * h * (cin + cout) + 9 * cin * cout)
            res[name] = {"ms": ms, "TFLOP/s": fl / ms / 1e9, "frac_tensor": fl / ms / 1e9 / PEAKS["tf"],
                         "GB/s_algorithmic": byts / ms / 1e6}
            ms = timeit(lambda: op.conv2d_tc(x, wp, 3), a.iters)
            res[name + " (plain epilogue)"] = {"ms": ms, "TFLOP/s": fl / ms / 1e9,
                                               "frac_tensor": fl / ms / 1e9 / PEAKS["tf"]}
            from transeditor_b200 import tc
            gy = torch.randn(b, cout, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: tc.wgrad_raw(gy, x, tc.Mode("s1", 3), (cout, cin, 3, 3)), a.iters)
            res[name.replace("conv2d_tc", "wgrad_tc")] = {"ms": ms, "TFLOP/s": fl / ms / 1e9,
                                                          "frac_tensor": fl / ms / 1e9 / PEAKS["tf"]}
            del gy
            del x
    for k, v in res.items():
        print(k, {kk: round(vv, 4) for kk, vv in v.items()}, flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"peaks": PEAKS, "results": res}, f, indent=1)


if __name__ == "__main__":
    main()
