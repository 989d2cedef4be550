"""Correctness + speed probe of the haloed-patch conv kernel (run once per TE_TC_HALO / TE_TC_HALO_BO setting):

    TE_TC_HALO=1 python tools/halo_probe.py
Prints per shape: max-abs error against an f32 torch convolution of the same bf16 operands, ms and TFLOP/s."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import tc  # noqa: E402

dev = "cuda"


def timeit(fn, iters=10):
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
    print("TE_TC_HALO=%s TE_TC_HALO_BO=%s" % (os.environ.get("TE_TC_HALO", "1"), os.environ.get("TE_TC_HALO_BO", "0")))
    g = torch.Generator().manual_seed(0)
    cases = [("s1", 16, 128, 128, 256, False), ("s1", 16, 256, 256, 128, False), ("s1", 16, 512, 512, 64, False),
             ("s1", 16, 128, 128, 256, True), ("s1", 16, 512, 512, 32, False), ("up", 16, 256, 128, 128, False),
             ("s1", 32, 128, 128, 256, False), ("s1", 2, 64, 128, 40, False)]
    for kind, b, cin, cout, h, per_sample in cases:
        x = torch.randn(b, cin, h, h, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if per_sample:
            w = (torch.randn(b, cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
        else:
            w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
        mode = tc.Mode("s1", 3) if kind == "s1" else tc.Mode("up", 3)
        wp = tc.pack_weight(w, False, 1.0)
        fn = lambda: tc.conv_raw(x, wp, mode)  # noqa: E731
        y = fn()
        # check a slice against torch (same bf16 operands, f32 math)
        nb = min(b, 2)
        xs, wq = x[:nb].float(), w.to(torch.bfloat16).float()
        if kind == "s1":
            if per_sample:
                ref = torch.cat([F.conv2d(xs[i:i + 1], wq[i], padding=1) for i in range(nb)])
            else:
                ref = F.conv2d(xs, wq, padding=1)
        else:
            ref = F.conv_transpose2d(xs, wq.transpose(0, 1), stride=2)
        err = (y[:nb].float() - ref).abs().max().item()
        last = (y[b - 1:].float() - (F.conv2d(x[b - 1:].float(), wq[b - 1] if per_sample else wq, padding=1) if kind == "s1"
                                     else F.conv_transpose2d(x[b - 1:].float(), wq.transpose(0, 1), stride=2))).abs().max().item()
        ms = timeit(fn)
        flops = 2.0 * b * h * h * cin * cout * 9
        print("%-3s b%-2d %3d->%3d @%3d ps%d  err %.3e / %.3e (ref max %.2f)  %.4f ms  %7.1f TFLOP/s"
              % (kind, b, cin, cout, h, per_sample, err, last, ref.abs().max().item(), ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
