"""Channels-last bf16 blur (upfirdn2d up=down=1) on the small tensors of the step: time and HBM fraction.

    TE_FIR_SMALL_TILES=0 python tools/fir_probe.py ; python tools/fir_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import op  # noqa: E402


def timeit(fn, iters=20):
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
    print("TE_FIR_SMALL_TILES=%s" % os.environ.get("TE_FIR_SMALL_TILES", "1"))
    k = torch.tensor([1., 3., 3., 1.])
    k = (k[None] * k[:, None] / 64).cuda()
    for b, c, h, pad in [(16, 512, 65, (1, 1)), (16, 512, 64, (2, 2)), (32, 512, 64, (2, 2)), (16, 512, 33, (1, 1)),
                         (16, 512, 17, (1, 1)), (16, 256, 129, (1, 1)), (16, 128, 257, (1, 1))]:
        x = torch.randn(b, c, h, h, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        fn = lambda: op.upfirdn2d(x, k, pad=pad)  # noqa: E731
        y = fn()
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x[:1, :8].float(), (pad[0], pad[1], pad[0], pad[1])),
                                         k.flip(0, 1)[None, None].repeat(8, 1, 1, 1), groups=8)
        err = (y[:1, :8].float() - ref).abs().max().item()
        ms = timeit(fn)
        byts = (x.numel() + y.numel()) * 2
        print("[%d,%d,%d,%d] pad%s -> %d  err %.2e  %.4f ms  %.0f GB/s (%.2f of 6555)"
              % (b, c, h, h, pad, y.shape[2], err, ms, byts / ms / 1e6, byts / ms / 1e6 / 6555), flush=True)


if __name__ == "__main__":
    main()
