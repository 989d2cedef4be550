"""Which N tile / CTA pairing is fastest for the MID and LOW resolution layers?  Run once per setting:
    TE_TC_BLOCK_N=128 TE_TC_2CTA=0 python tools/tc_tile_probe.py
Profiling aid (CUDA events, L2 flushed between launches)."""
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


tag = "N=%s 2cta=%s" % (os.environ.get("TE_TC_BLOCK_N", "auto"), os.environ.get("TE_TC_2CTA", "1"))
for (b, h, cin, cout, k) in ((16, 64, 512, 512, 3), (32, 64, 512, 512, 3), (16, 32, 512, 512, 3), (32, 32, 512, 512, 3),
                             (16, 16, 512, 512, 3), (32, 16, 512, 512, 3), (16, 8, 512, 512, 3), (16, 128, 256, 256, 3)):
    x = torch.randn(b, cin, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wp = op.pack_weight_tc(torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k))
    ms = timeit(lambda: op.conv2d_tc(x, wp, k))
    fl = 2.0 * b * h * h * cin * cout * k * k
    print("%s  b%d %d->%d @%d^2: %.1f us  %.0f TFLOP/s" % (tag, b, cin, cout, h, ms * 1e3, fl / ms / 1e9), flush=True)
