"""Standalone check of the tcgen05 path on a B200 (run under `timeout`; a descriptor mistake can hang).

    python tools/tc_selftest.py [--bench]
Prints max-abs errors of te_gemm_tc_selftest and te_conv2d_tc against torch references computed from
the same bf16-rounded operands in f32, and (with --bench) the time of BASELINE's flagship layers."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transeditor_b200 import lib, op  # noqa: E402

dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ok = True


def report(name, err, tol):
    global ok
    flag = "OK " if err <= tol else "BAD"
    if err > tol or not math.isfinite(err):
        ok = False
    print(f"[{flag}] {name}: max-abs err {err:.3e} (tol {tol:.1e})", flush=True)


def gemm_case(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16).to(dev)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16).to(dev)
    d = torch.full((m, n), float("nan"), device=dev)
    lib.gemm_tc_selftest(d, a, b, m, n, k)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    report(f"gemm {m}x{n}x{k}", (d - ref).abs().max().item(), 1e-3 * math.sqrt(k))


def conv_case(b, h, cin, cout, k, act=False, scale=False, bias=False, per_sample=False):
    g = torch.Generator().manual_seed(b * 1000 + h * 10 + cin + cout + k)
    x = torch.randn(b, cin, h, h, generator=g).to(torch.bfloat16).to(dev).contiguous(memory_format=torch.channels_last)
    if per_sample:
        w = (torch.randn(b, cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(dev)
    else:
        w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(dev)
    wp = op.pack_weight_tc(w)
    osc = (torch.rand(b, cout, generator=g) + 0.5).to(dev) if scale else None
    bi = torch.randn(cout, generator=g).to(dev) if bias else None
    y = op.conv2d_tc(x, wp, k, osc, bi, act)
    torch.cuda.synchronize()
    wb = w.to(torch.bfloat16).float()
    if per_sample:
        ref = torch.cat([F.conv2d(x[i:i + 1].float(), wb[i], padding=k // 2) for i in range(b)])
    else:
        ref = F.conv2d(x.float(), wb, padding=k // 2)
    if osc is not None:
        ref = ref * osc[:, :, None, None]
    if bi is not None:
        ref = ref + bi.view(1, -1, 1, 1)
    if act:
        ref = F.leaky_relu(ref, 0.2) * math.sqrt(2)
    err = (y.float() - ref).abs().max().item()
    report(f"conv b{b} {h}x{h} {cin}->{cout} k{k} act={act} scale={scale} bias={bias} ps={per_sample}", err,
           2e-2 * max(1.0, ref.abs().max().item() / 4))


def bench_case(b, h, cin, cout, k, iters=20):
    x = torch.randn(b, cin, h, h, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wp = op.pack_weight_tc(torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k))
    for _ in range(3):
        op.conv2d_tc(x, wp, k)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        op.conv2d_tc(x, wp, k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * b * h * h * cin * cout * k * k
    print(f"[bench] conv b{b} {h}x{h} {cin}->{cout} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0), flush=True)
    gemm_case(128, 64, 64)
    gemm_case(128, 128, 64)
    gemm_case(256, 128, 128)
    gemm_case(1024, 256, 512)
    conv_case(1, 16, 64, 64, 1)
    conv_case(1, 16, 64, 64, 3)
    conv_case(2, 16, 64, 128, 3)
    conv_case(3, 8, 128, 64, 3)
    conv_case(5, 4, 64, 64, 3)
    conv_case(2, 32, 128, 128, 3, act=True, scale=True, bias=True)
    conv_case(2, 64, 256, 192, 3, act=True, scale=True, bias=True)
    conv_case(2, 16, 64, 128, 3, per_sample=True, scale=True)
    conv_case(1, 128, 128, 128, 1, bias=True)
    if "--bench" in sys.argv:
        bench_case(16, 256, 128, 128, 3)
        bench_case(16, 128, 256, 256, 3)
        bench_case(16, 64, 512, 512, 3)
        bench_case(16, 32, 512, 512, 3)
    print("ALL OK" if ok else "FAILURES", flush=True)
    sys.exit(0 if ok else 1)
