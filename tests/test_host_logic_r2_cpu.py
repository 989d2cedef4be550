"""Host logic of the round-2 ops on CPU through tests/emu.py: the autograd formulas of tc.TcConvScaled, op.BlurBiasAct,
op.FromRGB, op.GroupedLinear / op.MappingColumns and the WeightEnergy kernels' wiring against plain torch autograd,
first and second order.  The kernels themselves are checked on the GPU (test_gpu_linear.py, test_gpu_from_rgb.py,
test_gpu_ops.py, test_gpu_reference_ops.py)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.usefixtures("cpu_emulation")


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("kind", ["s1", "up"])
def test_conv_scaled_gradients_first_and_second_order(kind):
    from transeditor_b200 import tc
    b, cin, cout, h = 2, 8, 16, 6
    x = _rand(b, cin, h, h, seed=1).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = _rand(cout, cin, 3, 3, seed=2, scale=0.2).requires_grad_(True)
    d = (torch.rand(b, cout, generator=torch.Generator().manual_seed(3)) + 0.5).requires_grad_(True)
    xr, wr, dr = (t.detach().double().requires_grad_(True) for t in (x, w, d))
    if kind == "s1":
        y = tc.conv2d_scaled(x, w, d, wscale=0.7)
        ref = F.conv2d(xr, wr * 0.7, padding=1) * dr[:, :, None, None]
    else:
        y = tc.conv_transpose2d_scaled(x, w, d, wscale=0.7)
        ref = F.conv_transpose2d(xr, (wr * 0.7).transpose(0, 1), stride=2) * dr[:, :, None, None]
    assert _rel(y, ref) < 1e-5
    gy = _rand(*y.shape, seed=4)
    g = torch.autograd.grad(y, (x, w, d), gy, create_graph=True)
    gr = torch.autograd.grad(ref, (xr, wr, dr), gy.double(), create_graph=True)
    for a, r in zip(g, gr):
        assert _rel(a, r) < 1e-4
    # second order: differentiate |dL/dx|^2 with respect to w and d
    gg = torch.autograd.grad(g[0].square().sum(), (w, d))
    ggr = torch.autograd.grad(gr[0].square().sum(), (wr, dr))
    for a, r in zip(gg, ggr):
        assert _rel(a, r) < 1e-3


def test_blur_bias_act_gradients():
    from transeditor_b200 import op
    x = _rand(2, 8, 9, 9, seed=5).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bias = _rand(8, seed=6, scale=0.3).requires_grad_(True)
    k = torch.tensor([1., 3., 3., 1.])
    fir = k[None] * k[:, None] / 16
    y = op.blur_bias_act(x, fir, (1, 1), bias)
    xr, br = x.detach().double().requires_grad_(True), bias.detach().double().requires_grad_(True)
    blur = F.conv2d(F.pad(xr, (1, 1, 1, 1)), fir.double().flip(0, 1)[None, None].repeat(8, 1, 1, 1), groups=8)
    ref = F.leaky_relu(blur + br.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert _rel(y, ref) < 1e-5
    gy = _rand(*y.shape, seed=7)
    g = torch.autograd.grad(y, (x, bias), gy)
    gr = torch.autograd.grad(ref, (xr, br), gy.double())
    for a, r in zip(g, gr):
        assert _rel(a, r) < 1e-4


def test_from_rgb_gradients_and_second_order_route():
    from transeditor_b200 import op
    img = (torch.rand(2, 3, 8, 8, generator=torch.Generator().manual_seed(8)) * 2 - 1).requires_grad_(True)
    w = _rand(16, 3, 1, 1, seed=9).requires_grad_(True)
    b = _rand(16, seed=10, scale=0.3).requires_grad_(True)
    y = op.from_rgb(img, w, b, 0.577, dtype=torch.float32)
    ir, wr, br = (t.detach().double().requires_grad_(True) for t in (img, w, b))
    ref = op.from_rgb_reference(ir, wr, br, 0.577, 2 ** 0.5)
    assert _rel(y, ref) < 1e-5
    gy = _rand(*y.shape, seed=11)
    g = torch.autograd.grad(y, (img, w, b), gy)
    gr = torch.autograd.grad(ref, (ir, wr, br), gy.double())
    for a, r in zip(g, gr):
        assert _rel(a.reshape(r.shape), r) < 1e-4
    # create_graph: the R1-style second order goes through the twice-differentiable tensor-core ops
    y2 = op.from_rgb(img, w, b, 0.577, dtype=torch.float32)
    (gi,) = torch.autograd.grad(y2.sum(), img, create_graph=True)
    (gw,) = torch.autograd.grad(gi.square().sum(), w)
    ref2 = op.from_rgb_reference(ir, wr, br, 0.577, 2 ** 0.5)
    (gir,) = torch.autograd.grad(ref2.sum(), ir, create_graph=True)
    (gwr,) = torch.autograd.grad(gir.square().sum(), wr)
    assert _rel(gw, gwr) < 1e-3


def test_grouped_linear_and_mapping_columns_formulas():
    from transeditor_b200 import op
    x = _rand(5, 24, seed=12).requires_grad_(True)
    w = _rand(10, 24, seed=13).requires_grad_(True)
    b = _rand(10, seed=14).requires_grad_(True)
    (y,) = op.grouped_linear([(x, w, b, 0.3, 0.5, True)])
    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = op._linear_composite(xr, wr, br, 0.3, 0.5, True, False)
    assert _rel(y, ref) < 1e-6
    gy = _rand(*y.shape, seed=15)
    for a, r in zip(torch.autograd.grad(y, (x, w, b), gy), torch.autograd.grad(ref, (xr, wr, br), gy.double())):
        assert _rel(a, r) < 1e-5
    code = _rand(3, 32, 16, seed=16).requires_grad_(True)
    ws = [(_rand(32, 32, seed=20 + i) * 3).requires_grad_(True) for i in range(12)]
    bs = [_rand(32, seed=40 + i).requires_grad_(True) for i in range(12)]
    out = op.mapping_columns(code, ws, bs, 0.05, 0.01)
    cr = code.detach().double().requires_grad_(True)
    wsr = [t.detach().double().requires_grad_(True) for t in ws]
    bsr = [t.detach().double().requires_grad_(True) for t in bs]
    ref = op.mapping_columns_reference(cr, wsr, bsr, 0.05, 0.01)
    assert _rel(out, ref) < 1e-5 and out[:, :, 12:].abs().sum() == 0
    gy = _rand(*out.shape, seed=17)
    got = torch.autograd.grad(out, [code] + ws + bs, gy)
    want = torch.autograd.grad(ref, [cr] + wsr + bsr, gy.double())
    for a, r in zip(got, want):
        assert _rel(a, r) < 1e-4


def test_weight_energy_matches_the_literal_form():
    from transeditor_b200.model import WeightEnergy
    w = _rand(6, 5, 3, 3, seed=18).requires_grad_(True)
    e = WeightEnergy.apply(w, 0.3)
    wr = w.detach().double().requires_grad_(True)
    ref = (wr * 0.3).pow(2).sum((2, 3))
    assert _rel(e, ref) < 1e-6
    g = _rand(6, 5, seed=19)
    (gw,) = torch.autograd.grad(e, w, g)
    (gwr,) = torch.autograd.grad(ref, wr, g.double())
    assert _rel(gw, gwr) < 1e-6
