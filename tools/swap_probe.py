"""Correctness + speed probe of the operand-swapped conv kernel (conv_tct.cuh), one process per setting:

    TE_TC_SWAP=0 python tools/swap_probe.py ; TE_TC_SWAP=1 python tools/swap_probe.py
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
    print("TE_TC_SWAP=%s TE_TC_DEBUG=%s" % (os.environ.get("TE_TC_SWAP", "1"), os.environ.get("TE_TC_DEBUG", "0")))
    g = torch.Generator().manual_seed(0)
    # kind, batch, cin, cout, size, k, per_sample, epilogue
    cases = [("s1", 16, 128, 128, 256, 3, False, False), ("s1", 16, 128, 128, 256, 3, True, False),
             ("s1", 32, 128, 128, 256, 3, False, False), ("s1", 8, 128, 128, 256, 3, True, False),
             ("s1", 16, 128, 128, 256, 3, False, True), ("up", 16, 256, 128, 128, 3, False, False),
             ("s1", 16, 64, 128, 256, 1, False, False), ("s1", 16, 8, 128, 256, 1, False, True),
             ("s1", 16, 256, 128, 128, 1, False, False), ("s1", 16, 512, 128, 40, 3, False, True),
             ("s1", 16, 256, 256, 128, 3, False, False), ("s1", 16, 128, 128, 256, 3, False, "act"),
             ("s1", 16, 128, 128, 256, 3, False, "res")]
    if os.environ.get("PROBE_CASES"):
        cases = [cases[int(i)] for i in os.environ["PROBE_CASES"].split(",")]
    for kind, b, cin, cout, h, k, per_sample, epi in cases:
        x = torch.randn(b, cin, h, h, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        shape = (b, cout, cin, k, k) if per_sample else (cout, cin, k, k)
        w = (torch.randn(*shape, generator=g) / (cin * k * k) ** 0.5).to(dev)
        mode = tc.Mode("s1", k) if kind == "s1" else tc.Mode("up", k)
        wp = tc.pack_weight(w, False, 1.0)
        kw = {}
        if epi == "res":
            kw["residual"] = torch.randn(b, cout, h, h, generator=g).to(dev).to(torch.bfloat16).contiguous(
                memory_format=torch.channels_last)
        elif epi:
            kw = dict(bias=torch.randn(cout, generator=g).to(dev), act=True)
            if epi is True:
                kw["out_scale"] = torch.rand(b, cout, generator=g).to(dev) + 0.5
            if kind == "s1" and epi is True:
                kw["residual"] = torch.randn(b, cout, h, h, generator=g).to(dev).to(torch.bfloat16).contiguous(
                    memory_format=torch.channels_last)
        fn = lambda: tc.conv_raw(x, wp, mode, **kw)  # noqa: E731
        y = fn()

        def ref_of(i):
            xs, wq = x[i:i + 1].float(), (w[i] if per_sample else w).to(torch.bfloat16).float()
            r = F.conv2d(xs, wq, padding=k // 2) if kind == "s1" else F.conv_transpose2d(xs, wq.transpose(0, 1), stride=2)
            if "out_scale" in kw:
                r = r * kw["out_scale"][i].view(1, -1, 1, 1)
            if "bias" in kw:
                r = F.leaky_relu(r + kw["bias"].view(1, -1, 1, 1), 0.2) * 2 ** 0.5
            if "residual" in kw:
                r = r + kw["residual"][i:i + 1].float()
            return r
        errs = [(y[i:i + 1].float() - ref_of(i)).abs().max().item() for i in (0, b - 1)]
        ms = timeit(fn)
        flops = 2.0 * b * h * h * cin * cout * k * k
        print("%-3s b%-2d %3d->%3d @%3d k%d ps%d epi%-5s  err %.3e / %.3e  %.4f ms  %7.1f TFLOP/s"
              % (kind, b, cin, cout, h, k, per_sample, str(epi), errs[0], errs[1], ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
