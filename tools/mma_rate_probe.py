"""MMA-only rate of the conv kernels by N tile (TE_TC_DEBUG=10: no loads, no epilogue work): is a tcgen05.mma's cost
fixed per instruction or proportional to N?  Prints cycles per k-block (4 MMAs of K16) per CTA (pair)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import tc  # noqa: E402

dev = "cuda"


def run(b, cin, cout, h):
    x = torch.randn(b, cin, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, 3, 3, device=dev)
    wp = tc.pack_weight(w, False, 1.0)
    mode = tc.Mode("s1", 3)
    for _ in range(3):
        tc.conv_raw(x, wp, mode)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tc.conv_raw(x, wp, mode)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * b * h * h * cin * cout * 9
    return ms, flops / ms / 1e9


print("TE_TC_DEBUG=%s TE_TC_2CTA=%s TE_TC_HALO=%s TE_TC_BLOCK_N=%s" % tuple(os.environ.get(k, "-") for k in
      ("TE_TC_DEBUG", "TE_TC_2CTA", "TE_TC_HALO", "TE_TC_BLOCK_N")))
for cout in (64, 128, 256, 512):
    ms, tf = run(16, 128, cout, 256 if cout <= 128 else 128)
    print("128->%3d  %.4f ms  %7.1f TFLOP/s" % (cout, ms, tf), flush=True)
