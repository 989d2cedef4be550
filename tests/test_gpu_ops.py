"""Parity of every C-ABI kernel with the CPU oracle on seeded inputs (B200 box)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops_cpu
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)


def _tol(dtype):
    # alpha / scale cross the C ABI as `float` (as in the reference's fused_bias_act signature), so
    # f64 results carry the f32 rounding of sqrt(2): ~2e-8 relative
    return {torch.float64: 2e-7, torch.float32: 2e-6, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dtype]


# ------------------------------------------------------------------------------- fused_bias_act
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,cl", [((3, 5, 4, 6), False), ((2, 8, 16, 16), False), ((2, 8, 16, 16), True),
                                      ((7, 512), False), ((4, 6, 5, 5), True), ((1, 3, 1, 1), False),
                                      ((2, 16, 65, 33), False)])
def test_fused_leaky_relu_forward(dtype, shape, cl):
    from utils.op import fused_leaky_relu
    x = _rand(*shape, dtype=dtype, seed=1)
    b = _rand(shape[1], dtype=dtype, seed=2)
    xd = x.to(DEV)
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    y = fused_leaky_relu(xd, b.to(DEV))
    ref = ops_cpu.fused_leaky_relu(x.double(), b.double())
    assert y.shape == x.shape and y.dtype == dtype
    if cl:
        assert y.is_contiguous(memory_format=torch.channels_last)
    err = (y.double().cpu() - ref).abs().max().item()
    assert err <= _tol(dtype) * max(1.0, ref.abs().max().item())


def test_fused_bias_act_all_codes_golden():
    """Every (act, grad) case of the native entry point against the golden vectors."""
    from transeditor_b200 import lib
    gold = load_golden("ops")
    x = torch.from_numpy(gold["fba_x"]).to(DEV)
    b = torch.from_numpy(gold["fba_b"]).to(DEV)
    r = torch.from_numpy(gold["fba_ref"]).to(DEV)
    hw = x.shape[2] * x.shape[3]

    def run(bias, ref, act, grad):
        out = torch.empty_like(x)
        lib.fused_bias_act(out, x, bias, ref, act, grad, 0.2, 2 ** 0.5, hw, x.shape[1])
        return out.cpu().numpy()

    assert np.abs(run(b, None, 3, 0) - gold["fba_30"]).max() < 1e-6
    assert np.abs(run(b, r, 3, 1) - gold["fba_31"]).max() < 1e-6
    assert np.abs(run(None, r, 3, 1) - gold["fba_31_nobias"]).max() < 1e-6
    assert np.abs(run(b, None, 1, 0) - gold["fba_10"]).max() < 1e-6
    assert np.abs(run(b, r, 3, 2) - gold["fba_32"]).max() == 0.0
    with pytest.raises(RuntimeError, match="act must be"):
        run(b, None, 2, 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape,cl", [((3, 5, 4, 6), False), ((2, 8, 16, 16), True), ((7, 512), False),
                                      ((2, 4, 70, 70), False), ((9, 12), False), ((3, 128, 9, 9), True),
                                      ((2, 520, 4, 4), True)])
def test_fused_leaky_relu_first_and_second_order(dtype, shape, cl):
    from utils.op import fused_leaky_relu
    x = _rand(*shape, dtype=dtype, seed=3)
    b = _rand(shape[1], dtype=dtype, seed=4)
    gy = _rand(*shape, dtype=dtype, seed=5)
    ggx = _rand(*shape, dtype=dtype, seed=6)
    ggb = _rand(shape[1], dtype=dtype, seed=7)

    def run(fn, dev, fmt):
        xx = x.to(dev)
        if fmt:
            xx = xx.contiguous(memory_format=torch.channels_last)
        xx = xx.requires_grad_(True)
        bb = b.to(dev).requires_grad_(True)
        y = fn(xx, bb)
        gx, gb = torch.autograd.grad(y, (xx, bb), gy.to(dev), create_graph=True)
        # second order: only grad_output receives a gradient (sign comes from the saved output)
        gyy = gy.to(dev).requires_grad_(True)
        gx2, gb2 = torch.autograd.grad(fn(xx, bb), (xx, bb), gyy, create_graph=True)
        (gg,) = torch.autograd.grad((gx2 * ggx.to(dev)).sum() + (gb2 * ggb.to(dev)).sum(), gyy)
        return [t.detach().double().cpu() for t in (y, gx, gb, gg)]

    got = run(fused_leaky_relu, DEV, cl)
    ref = run(ops_cpu.fused_leaky_relu, "cpu", False)
    for a, r, nm in zip(got, ref, ("y", "gx", "gb", "gg")):
        tol = (2e-7 if dtype == torch.float64 else 3e-5) * max(1.0, r.abs().max().item())
        assert (a - r).abs().max().item() <= tol, nm


def test_fused_bias_act_large_property():
    """BASELINE size [16,128,256,256]: y(-x) relation and exact scale of positive entries."""
    from utils.op import fused_leaky_relu
    x = torch.randn(16, 128, 256, 256, device=DEV)
    b = torch.zeros(128, device=DEV)
    y = fused_leaky_relu(x, b)
    yn = fused_leaky_relu(-x, b)
    # lrelu(x) - lrelu(-x) * ... : y - (-yn) = sqrt2 * 1.2 * x  for every element
    assert torch.allclose(y - yn, x * (math.sqrt(2) * 1.2), rtol=1e-5, atol=1e-6)
    del y, yn, x


# ------------------------------------------------------------------------------- upfirdn2d
def test_upfirdn2d_golden_all_parameter_sets():
    """f64 through the public API against the kernel-index transcription goldens (asymmetric FIR)."""
    from utils.op import upfirdn2d
    gold = load_golden("ops")
    fir = torch.from_numpy(gold["fir"]).to(DEV)
    for ci, (up, down, p0, p1) in enumerate(gold["cases"]):
        for hw in ("5x7", "8x8", "9x6"):
            x = torch.from_numpy(gold[f"up_x_{ci}_{hw}"])[None, None].to(DEV)
            y = upfirdn2d(x, fir, up=int(up), down=int(down), pad=(int(p0), int(p1)))
            ref = gold[f"up_y_{ci}_{hw}"]
            assert tuple(y.shape[2:]) == ref.shape
            # FIR is passed to the kernel in f32: error bounded by f32 rounding of the taps
            assert np.abs(y[0, 0].cpu().numpy() - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("case", [
    # (n, c, h, w, up, down, pad)  — hot tiled path (up=down=1, planes) and generic path
    (2, 3, 40, 70, 1, 1, (1, 1)), (2, 3, 40, 70, 1, 1, (2, 2)), (1, 2, 129, 129, 1, 1, (1, 1)),
    (1, 2, 257, 257, 1, 1, (1, 1)), (1, 1, 33, 200, 1, 1, (2, 1)), (2, 3, 16, 16, 2, 1, (2, 1)),
    (2, 3, 32, 32, 1, 2, (1, 1)), (1, 4, 9, 9, 1, 1, (2, 2)), (1, 2, 64, 64, 1, 1, (0, 0)),
    (1, 2, 40, 36, 1, 1, (-1, 2)), (2, 16, 20, 24, 1, 1, (1, 1)), (1, 8, 33, 17, 1, 1, (2, 2)),
    (2, 128, 9, 9, 1, 1, (2, 2)), (1, 24, 16, 16, 1, 1, (1, 1)),
])
@pytest.mark.parametrize("cl", [False, True])
def test_upfirdn2d_matches_oracle(dtype, case, cl):
    from utils.op import upfirdn2d
    n, c, h, w, up, down, pad = case
    x = _rand(n, c, h, w, dtype=dtype, seed=11)
    fir = torch.tensor([1., 3., 3., 1.])
    fir = torch.outer(fir, fir)
    fir = fir / fir.sum() * (up * up)
    xd = x.to(DEV)
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    y = upfirdn2d(xd, fir.to(DEV), up=up, down=down, pad=pad)
    ref = ops_cpu.upfirdn2d(x.double(), fir.double(), up, down, pad)
    assert y.shape == ref.shape and y.dtype == dtype
    tol = {torch.float64: 1e-7, torch.float32: 2e-6, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dtype]
    assert (y.double().cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("case", [
    # (n, c, h, w, pad, fir) — large channels-last inputs take the TMA-staged kernel (>= 2^20 outputs, c % 64 == 0;
    # f32: 32-channel chunks, the fp32 parity mode's layout)
    (2, 128, 65, 67, (1, 1), "sep"), (1, 64, 129, 129, (2, 2), "sep"), (1, 192, 80, 96, (2, 1), "sep"),
    (2, 64, 97, 100, (-1, 2), "sep"), (1, 128, 257, 36, (1, 1), "sep"), (3, 64, 90, 70, (2, 2), "full"),
    (1, 128, 100, 90, (1, 1), "k3"), (1, 64, 300, 60, (0, 0), "sep"), (1, 64, 20, 900, (2, 2), "sep"),
])
def test_upfirdn2d_tma_staged_channels_last(dtype, case):
    from utils.op import upfirdn2d
    n, c, h, w, pad, kind = case
    x = _rand(n, c, h, w, dtype=dtype, seed=31)
    if kind == "sep":
        fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])) / 64
    elif kind == "full":
        fir = _rand(4, 4, seed=32).abs() / 8
    else:
        fir = _rand(3, 3, seed=33).abs() / 4
    y = upfirdn2d(x.to(DEV).contiguous(memory_format=torch.channels_last), fir.to(DEV), pad=pad)
    assert y.is_contiguous(memory_format=torch.channels_last)
    ref = ops_cpu.upfirdn2d(x.double(), fir.double(), 1, 1, pad)
    assert y.shape == ref.shape and y.numel() >= 1 << 20
    tol = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2, torch.float32: 2e-6}[dtype]
    assert (y.double().cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("case", [
    # (n, c, h, w, up, down, pad, fir): channels-last factor-2 resampling on the TMA-staged kernels
    (2, 64, 64, 64, 1, 2, (1, 1), "sep"), (1, 128, 33, 70, 1, 2, (1, 1), "sep"), (1, 64, 40, 36, 1, 2, (2, 2), "sep"),
    (2, 64, 40, 40, 1, 2, (0, 0), "full"), (1, 64, 90, 50, 1, 2, (-1, 2), "sep"), (4, 64, 16, 16, 1, 2, (1, 1), "sep"),
    (2, 64, 32, 32, 2, 1, (2, 1), "sep"), (1, 64, 31, 45, 2, 1, (2, 1), "sep"), (1, 128, 40, 40, 2, 1, (1, 1), "sep"),
    (1, 64, 40, 36, 2, 1, (3, 2), "full"), (1, 64, 30, 30, 2, 1, (0, 0), "sep"), (1, 64, 25, 25, 2, 1, (-1, 3), "k3"),
    (8, 64, 8, 8, 2, 1, (2, 1), "sep"), (1, 192, 64, 64, 1, 2, (1, 1), "k3"),
])
def test_upfirdn2d_tma_resample_channels_last(dtype, case):
    from utils.op import upfirdn2d
    n, c, h, w, up, down, pad, kind = case
    x = _rand(n, c, h, w, dtype=dtype, seed=41)
    if kind == "sep":
        fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])) / 64 * up * up
    elif kind == "full":
        fir = _rand(4, 4, seed=42).abs() / 8 * up * up
    else:
        fir = _rand(3, 3, seed=43).abs() / 4 * up * up
    y = upfirdn2d(x.to(DEV).contiguous(memory_format=torch.channels_last), fir.to(DEV), up=up, down=down, pad=pad)
    ref = ops_cpu.upfirdn2d(x.double(), fir.double(), up, down, pad)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    tol = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2, torch.float32: 2e-6}[dtype]
    assert (y.double().cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


def test_upfirdn2d_down2_gradient_is_the_up2_kernel():
    """Backward of the decimating FIR (ResBlock skip) = up=2 FIR of the compact gradient; both TMA kernels."""
    from utils.op import upfirdn2d
    x = _rand(2, 64, 64, 64, dtype=torch.bfloat16, seed=45)
    fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.])) / 64
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = upfirdn2d(xd, fir.to(DEV), down=2, pad=(1, 1))
    g = _rand(*y.shape, dtype=torch.bfloat16, seed=46)
    (gx,) = torch.autograd.grad(y, xd, g.to(DEV).contiguous(memory_format=torch.channels_last))
    xr = x.double().requires_grad_(True)
    (gr,) = torch.autograd.grad(ops_cpu.upfirdn2d(xr, fir.double(), 1, 2, (1, 1)), xr, g.double())
    assert (gx.double().cpu() - gr).abs().max().item() <= 1.6e-2 * max(1.0, gr.abs().max().item())


def test_upfirdn2d_channels_last_nonseparable_fir():
    """The vectorised channels-last kernel takes a separable fast path; a random (rank-4) FIR must go
    through its general path and still match."""
    from utils.op import upfirdn2d
    x = _rand(2, 16, 21, 19, seed=16)
    fir = _rand(4, 4, seed=17).abs()
    y = upfirdn2d(x.to(DEV).contiguous(memory_format=torch.channels_last), fir.to(DEV), pad=(2, 1))
    ref = ops_cpu.upfirdn2d(x.double(), fir.double(), 1, 1, (2, 1))
    assert (y.double().cpu() - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    fir3 = _rand(3, 3, seed=18).abs()
    y = upfirdn2d(x.to(DEV).contiguous(memory_format=torch.channels_last), fir3.to(DEV), pad=(1, 1))
    ref = ops_cpu.upfirdn2d(x.double(), fir3.double(), 1, 1, (1, 1))
    assert (y.double().cpu() - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(2, 3, 40, 70, 1, 1, (1, 1)), (2, 3, 12, 12, 2, 1, (2, 1)),
                                  (1, 2, 34, 34, 1, 1, (2, 2)), (2, 2, 16, 16, 1, 2, (1, 1))])
def test_upfirdn2d_first_and_second_order(case):
    from utils.op import upfirdn2d
    n, c, h, w, up, down, pad = case
    dtype = torch.float64
    x = _rand(n, c, h, w, dtype=dtype, seed=12)
    fir = _rand(4, 4, dtype=dtype, seed=13).abs()

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        y = fn(xx, fir.to(dev), up, down, pad)
        gy = _rand(*y.shape, dtype=dtype, seed=14).to(dev).requires_grad_(True)
        (gx,) = torch.autograd.grad(y, xx, gy, create_graph=True)
        ggx = _rand(*gx.shape, dtype=dtype, seed=15).to(dev)
        (gg,) = torch.autograd.grad(gx, gy, ggx)
        return [t.detach().cpu() for t in (y, gx, gg)]

    got = run(upfirdn2d, DEV)
    ref = run(ops_cpu.upfirdn2d, "cpu")
    for a, r in zip(got, ref):
        assert (a - r).abs().max().item() < 1e-6 * max(1.0, r.abs().max().item())


def test_upfirdn2d_full_size_properties():
    """BASELINE flagship [16,128,257,257] -> 256^2 (1.08 GB): linearity and DC gain (sum of FIR = 1)."""
    from utils.op import upfirdn2d
    fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.]))
    fir = (fir / fir.sum()).to(DEV)
    a = torch.randn(16, 128, 257, 257, device=DEV)
    ya = upfirdn2d(a, fir, pad=(1, 1))
    assert ya.shape == (16, 128, 256, 256)
    # interior of a constant plane is reproduced exactly up to rounding
    a.fill_(1.5)
    yc = upfirdn2d(a, fir, pad=(1, 1))
    assert (yc[:, :, 2:-2, 2:-2] - 1.5).abs().max().item() < 1e-6
    # linearity on a slice (keeps memory bounded)
    b1 = torch.randn(2, 128, 257, 257, device=DEV)
    b2 = torch.randn(2, 128, 257, 257, device=DEV)
    lhs = upfirdn2d(2.0 * b1 - 3.0 * b2, fir, pad=(1, 1))
    rhs = 2.0 * upfirdn2d(b1, fir, pad=(1, 1)) - 3.0 * upfirdn2d(b2, fir, pad=(1, 1))
    assert (lhs - rhs).abs().max().item() < 1e-5
    # spot-check one plane against the oracle
    ref = ops_cpu.upfirdn2d(b1[:1, :2].double().cpu(), fir.double().cpu(), 1, 1, (1, 1))
    assert (upfirdn2d(b1, fir, pad=(1, 1))[:1, :2].double().cpu() - ref).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------- convolution
CONV_CASES = [
    # b, cin, cout, h, w, k, kind   kind: plain(stride,pad) or transposed
    (2, 8, 5, 9, 7, 3, ("conv", 1, 1)), (2, 70, 66, 12, 12, 3, ("conv", 1, 1)),
    (3, 16, 3, 8, 8, 1, ("conv", 1, 0)), (2, 6, 9, 11, 11, 3, ("conv", 2, 0)),
    (2, 6, 9, 12, 10, 1, ("conv", 2, 0)), (2, 8, 5, 6, 5, 3, ("convT", 2, 0)),
    (1, 130, 20, 17, 17, 3, ("conv", 1, 1)), (2, 3, 128, 16, 16, 1, ("conv", 1, 0)),
    (1, 33, 40, 4, 4, 3, ("conv", 1, 1)),
]


def _ref_conv(x, w, kind):
    if kind[0] == "conv":
        return F.conv2d(x, w, stride=kind[1], padding=kind[2])
    return F.conv_transpose2d(x, w.transpose(0, 1), stride=kind[1], padding=0)


def _our_conv(x, w, kind):
    from transeditor_b200 import op
    if kind[0] == "conv":
        return op.conv2d(x, w, stride=kind[1], padding=kind[2])
    return op.conv_transpose2d(x, w, stride=kind[1])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_backward_double_backward(dtype, case):
    b, cin, cout, h, w_, k, kind = case
    x = _rand(b, cin, h, w_, dtype=dtype, seed=21)
    w = _rand(cout, cin, k, k, dtype=dtype, seed=22) / math.sqrt(cin * k * k)

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        ww = w.to(dev).requires_grad_(True)
        y = fn(xx, ww, kind)
        gy = _rand(*y.shape, dtype=dtype, seed=23).to(dev).requires_grad_(True)
        gx, gw = torch.autograd.grad(y, (xx, ww), gy, create_graph=True)
        ggx = _rand(*gx.shape, dtype=dtype, seed=24).to(dev)
        ggw = _rand(*gw.shape, dtype=dtype, seed=25).to(dev)
        # second order w.r.t. everything (R1 / path-length use all of these routes)
        s = (gx * ggx).sum() + (gw * ggw).sum()
        h_gy, h_x, h_w = torch.autograd.grad(s, (gy, xx, ww))
        return [t.detach().double().cpu() for t in (y, gx, gw, h_gy, h_x, h_w)]

    got = run(_our_conv, DEV)
    ref = run(_ref_conv, "cpu")
    for a, r, nm in zip(got, ref, ("y", "gx", "gw", "h_gy", "h_x", "h_w")):
        assert a.shape == r.shape, nm
        tol = (1e-10 if dtype == torch.float64 else 5e-5) * max(1.0, r.abs().max().item())
        assert (a - r).abs().max().item() <= tol, (nm, (a - r).abs().max().item())


def test_conv_fused_epilogue_matches_composition():
    """in_scale / out_scale / noise / bias / activation inside the kernel == the unfused composition."""
    from transeditor_b200 import op
    b, cin, cout, h = 3, 24, 40, 10
    x = _rand(b, cin, h, h, seed=31).to(DEV)
    w = (_rand(cout, cin, 3, 3, seed=32) / 15).to(DEV)
    s = (_rand(b, cin, seed=33).abs() + 0.5).to(DEV)
    d = (_rand(b, cout, seed=34).abs() + 0.5).to(DEV)
    bias = _rand(cout, seed=35).to(DEV)
    noise = _rand(b, 1, h, h, seed=36).to(DEV)
    nw = torch.tensor([0.3], device=DEV)
    y = op.conv2d_fused(x, w, in_scale=s, out_scale=d, bias=bias, noise=noise, noise_w=nw, act=True, padding=1)
    ref = F.conv2d((x * s[:, :, None, None]).double().cpu(), w.double().cpu(), padding=1) * d.double().cpu()[:, :, None, None]
    ref = ref + 0.3 * noise.double().cpu() + bias.double().cpu().view(1, -1, 1, 1)
    ref = F.leaky_relu(ref, 0.2) * math.sqrt(2)
    assert (y.double().cpu() - ref).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item())
    # shared noise map ([1,1,H,W]) and transposed form
    y2 = op.conv2d_fused(x, w, in_scale=s, out_scale=d, transpose_stride=2)
    ref2 = F.conv_transpose2d((x * s[:, :, None, None]).double().cpu(), w.double().cpu().transpose(0, 1), stride=2)
    ref2 = ref2 * d.double().cpu()[:, :, None, None]
    assert (y2.double().cpu() - ref2).abs().max().item() < 5e-5 * max(1.0, ref2.abs().max().item())


def test_conv_gradcheck_f64():
    from transeditor_b200 import op
    x = _rand(2, 3, 5, 5, dtype=torch.float64, seed=41).to(DEV).requires_grad_(True)
    w = _rand(4, 3, 3, 3, dtype=torch.float64, seed=42).to(DEV).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, b: op.conv2d(a, b, 1, 1), (x, w), eps=1e-6, atol=1e-6)
    assert torch.autograd.gradgradcheck(lambda a, b: op.conv2d(a, b, 2, 0), (x, w), eps=1e-6, atol=1e-6)
    assert torch.autograd.gradcheck(lambda a, b: op.conv_transpose2d(a, b, 2), (x, w), eps=1e-6, atol=1e-6)
    assert torch.autograd.gradgradcheck(lambda a, b: op.conv_transpose2d(a, b, 2), (x, w), eps=1e-6, atol=1e-6)


def test_conv_flagship_shape_spotcheck():
    """convs.11 of BASELINE cfg: 128->128 @256^2, B=16 (309 GFLOP): shape + one sample's rows vs oracle."""
    from transeditor_b200 import op
    x = torch.randn(16, 128, 256, 256, device=DEV)
    w = torch.randn(128, 128, 3, 3, device=DEV) / math.sqrt(128 * 9)
    y = op.conv2d_fused(x, w, padding=1)
    assert y.shape == (16, 128, 256, 256)
    ref = F.conv2d(x[5:6, :, 100:110].double().cpu(), w.double().cpu(), padding=(0, 1))
    assert (y[5:6, :, 101:109].double().cpu() - ref).abs().max().item() < 2e-4
    # linearity in x (size-independent property)
    x2 = torch.randn(2, 128, 256, 256, device=DEV)
    lhs = op.conv2d_fused(2 * x[:2] - x2, w, padding=1)
    rhs = 2 * y[:2] - op.conv2d_fused(x2, w, padding=1)
    assert (lhs - rhs).abs().max().item() < 1e-3


# ------------------------------------------------------------------------------- attention / adam
def test_attention_core_matches_reference_formula():
    from transeditor_b200 import op
    q, k, v = (_rand(5, 16, 128, seed=s).to(DEV) for s in (51, 52, 53))
    out, sim = op.attn_core(q, k, v)
    # literal reference form (model_spatial_query.py:888-894) in f64 on the CPU
    qq, kk, vv = (t.double().cpu() for t in (q, k, v))
    n, l, g, gp = 5, 16, 4, 32
    qh = qq.reshape(n, l, g, gp).permute(0, 2, 3, 1)
    kh = kk.reshape(n, l, g, gp).permute(0, 2, 3, 1)
    vh = vv.reshape(n, l, g, gp).permute(0, 2, 3, 1)
    s = torch.softmax(torch.einsum("abcd,abce->abde", qh, kh) * 128 ** -0.5, dim=3)
    sv = torch.einsum("abcd,abed->abec", s, vh).reshape(n, 128, l).permute(0, 2, 1)
    assert (sim.double().cpu() - s).abs().max().item() < 1e-5
    assert (out.double().cpu() - sv).abs().max().item() < 1e-5
    # gradients of the fused op == gradients of the literal form
    q2, k2, v2 = (t.clone().requires_grad_(True) for t in (q, k, v))
    o2, _ = op.attn_core(q2, k2, v2)
    go = _rand(5, 16, 128, seed=54).to(DEV)
    grads = torch.autograd.grad(o2, (q2, k2, v2), go)
    q3, k3, v3 = (t.clone().requires_grad_(True) for t in (q, k, v))
    o3, _ = op.attn_core_reference(q3, k3, v3)
    ref = torch.autograd.grad(o3, (q3, k3, v3), go)
    for a, r in zip(grads, ref):
        assert (a - r).abs().max().item() < 1e-5


def test_adam_ema_matches_torch_adam():
    from transeditor_b200 import lib
    n = 10007
    p0 = _rand(n, seed=61).to(DEV)
    p = p0.clone()
    ema = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    q = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([q], lr=0.0016, betas=(0.0, 0.9919), eps=1e-8)
    ema_ref = p0.clone()
    for step in range(1, 6):
        g = _rand(n, seed=70 + step).to(DEV)
        lib.adam_ema(p, g * 2.0, m, v, ema, 0.0016, 0.0, 0.9919, 1e-8, step, 0.998, 0.5)
        q.grad = g.clone()
        opt.step()
        ema_ref.mul_(0.998).add_(q.detach(), alpha=0.002)
    assert (p - q.detach()).abs().max().item() < 5e-6
    assert (ema - ema_ref).abs().max().item() < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,pad", [((4, 128, 129, 129), (1, 1)), ((2, 64, 66, 40), (2, 2)), ((16, 128, 257, 257), (1, 1)),
                                       ((2, 24, 33, 33), (1, 1))])
def test_blur_bias_act_matches_the_two_step_form(dtype, shape, pad):
    """op.blur_bias_act (te_upfirdn2d_bias_act, or the two-launch route for geometries the fused kernel does not cover)
    == fused_leaky_relu(upfirdn2d(x)) of the existing operators; gradients and double backward included."""
    from transeditor_b200 import op
    g = torch.Generator().manual_seed(3)
    n, c, h, w = shape
    k = torch.tensor([1., 3., 3., 1.])
    fir = (k[None] * k[:, None] / 16).cuda()

    def make():
        x = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(1)).cuda().to(dtype)
        x = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        b = (torch.randn(c, generator=torch.Generator().manual_seed(2)) * 0.5).cuda().requires_grad_(True)
        return x, b
    x1, b1 = make()
    y1 = op.blur_bias_act(x1, fir, pad, b1)
    x2, b2 = make()
    y2 = op.fused_leaky_relu(op.upfirdn2d(x2, fir, pad=pad), b2)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert y1.shape == y2.shape and y1.dtype == dtype
    assert (y1.float() - y2.float()).abs().max().item() <= tol * max(1.0, y2.float().abs().max().item())
    gy = torch.randn(y2.shape, generator=g).cuda().to(dtype)
    (gx1, gb1) = torch.autograd.grad(y1, (x1, b1), gy, create_graph=True)
    (gx2, gb2) = torch.autograd.grad(y2, (x2, b2), gy, create_graph=True)
    if dtype == torch.float32:
        assert (gx1 - gx2).abs().max().item() <= tol * max(1.0, gx2.abs().max().item())
        assert (gb1 - gb2).abs().max().item() <= 1e-3 * max(1.0, gb2.abs().max().item())
    else:
        # the two forms round the pre-activation differently, so the leaky-ReLU mask of a few near-zero elements flips
        cos = torch.nn.functional.cosine_similarity(gx1.float().flatten(), gx2.float().flatten(), dim=0).item()
        assert cos > 0.999
        assert (gb1.float() - gb2.float()).abs().max().item() <= 5e-2 * max(1.0, gb2.float().abs().max().item())
    if dtype == torch.float32 and n * c * h * w < 4_000_000:
        (gg1,) = torch.autograd.grad(gx1.square().sum(), x1, allow_unused=True)
        (gg2,) = torch.autograd.grad(gx2.square().sum(), x2, allow_unused=True)
        assert (gg1 is None) == (gg2 is None)
