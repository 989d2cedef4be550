"""Speed + correctness probe of the tcgen05 weight-gradient kernel per cin-tile width (one process per setting):

    TE_WG_N=128 python tools/wgrad_probe.py ; python tools/wgrad_probe.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import tc  # noqa: E402


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
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
    print("TE_WG_N=%s" % os.environ.get("TE_WG_N", "auto"))
    g = torch.Generator().manual_seed(0)
    cases = [(16, 512, 512, 64, False), (16, 256, 256, 128, False), (16, 128, 128, 256, False), (16, 512, 512, 32, False),
             (8, 256, 256, 128, True), (16, 256, 256, 128, True), (16, 128, 128, 256, True), (8, 128, 128, 256, True),
             (16, 256, 512, 64, False), (16, 512, 512, 16, False), (2, 256, 128, 24, False)]
    for b, cin, cout, h, ps in cases:
        x = torch.randn(b, cin, h, h, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gy = torch.randn(b, cout, h, h, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        shape = (b, cout, cin, 3, 3) if ps else (cout, cin, 3, 3)
        fn = lambda: tc.wgrad_raw(gy, x, tc.Mode("s1", 3), shape)  # noqa: E731
        gw = fn()
        nb = 1 if ps else min(b, 2)
        xs = x[:nb].float().requires_grad_(False)
        w0 = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
        xall = x.float() if not ps else x[:1].float()
        gall = gy.float() if not ps else gy[:1].float()
        (ref,) = torch.autograd.grad(F.conv2d(xall, w0, padding=1), w0, gall)
        got = gw[0] if ps else gw
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        ms = timeit(fn)
        fl = 2.0 * b * h * h * cin * cout * 9
        print("b%-2d %3d->%3d @%3d ps%d  rel err %.2e  %.4f ms  %7.1f TFLOP/s (incl. zero fill + unpack)"
              % (b, cin, cout, h, ps, err, ms, fl / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
