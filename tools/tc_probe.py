"""Where does conv_tc_kernel's tile time go?  Times two shapes with TE_TC_DEBUG variants
(1: no global stores, 2: no epilogue work at all, 4: no MMAs); run once per variant:
    TE_TC_DEBUG=0 python tools/tc_probe.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import op  # noqa: E402

dev = "cuda"


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for (b, h, cin, cout, k) in ((16, 256, 128, 128, 3), (16, 256, 8, 128, 1), (16, 128, 256, 256, 3), (16, 64, 512, 512, 3)):
    x = torch.randn(b, cin, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wp = op.pack_weight_tc(torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k))
    ms = timeit(lambda: op.conv2d_tc(x, wp, k))
    tiles = b * h * h / 128 * (cout / 128)
    print("TE_TC_DEBUG=%s  %d->%d k%d @%d: %.3f ms  (%.2f us per tile per SM, %d tiles)" % (
        os.environ.get("TE_TC_DEBUG", "0"), cin, cout, k, h, ms, ms * 1e3 / (tiles / 148), tiles), flush=True)
