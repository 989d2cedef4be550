"""te_linear_grouped / te_linear_wgrad_grouped (grouped EqualLinear, mapping columns) against float64 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _layers(specs, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for m, k, n, bias, act, alpha, bmul in specs:
        x = torch.randn(m, k, generator=g).cuda().requires_grad_(True)
        w = torch.randn(n, k, generator=g).cuda().requires_grad_(True)
        b = torch.randn(n, generator=g).cuda().requires_grad_(True) if bias else None
        out.append((x, w, b, alpha, bmul, act))
    return out


SPECS = [(16, 512, 512, True, True, 0.0442, 0.01), (16, 512, 128, True, False, 0.0442, 1.0), (8, 512, 3, True, False, 1.0, 1.0),
         (1, 512, 512, True, False, 0.5, 1.0), (35, 512, 256, False, True, 0.1, 1.0), (16, 16, 14, True, False, 0.25, 1.0),
         (5, 40, 24, True, True, 1.0, 2.0), (32, 512, 1, True, False, 0.0442, 1.0), (64, 528, 136, True, False, 0.3, 1.0)]


@pytest.mark.parametrize("tf32,tol", [(False, 3e-5), (True, 2e-2)])
def test_grouped_linear_forward_and_gradients(tf32, tol):
    from transeditor_b200 import op
    layers = _layers(SPECS)
    ys = op.grouped_linear(layers, tf32=tf32)
    loss, ref_loss = 0, 0
    gs = [torch.randn(y.shape, generator=torch.Generator().manual_seed(9 + i)).cuda() for i, y in enumerate(ys)]
    refs = []
    for (x, w, b, alpha, bmul, act), y, g in zip(layers, ys, gs):
        xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
        bd = b.detach().double().requires_grad_(True) if b is not None else None
        r = op._linear_composite(xd, wd, bd, alpha, bmul, act, False)
        assert _rel(y, r) < tol
        refs.append((xd, wd, bd))
        loss = loss + (y * g).sum()
        ref_loss = ref_loss + (r * g.double()).sum()
    loss.backward()
    ref_loss.backward()
    for (x, w, b, _, _, act), (xd, wd, bd) in zip(layers, refs):
        if tf32 and act:
            continue   # single-pass TF32 flips the sign of outputs next to zero, i.e. the leaky-ReLU mask of a few elements
        assert _rel(x.grad, xd.grad) < tol
        assert _rel(w.grad, wd.grad) < max(tol, 2e-6)
        if b is not None:
            assert _rel(b.grad, bd.grad) < max(tol, 2e-6)


def test_grouped_linear_strided_inputs_and_split_k():
    from transeditor_b200 import op
    g = torch.Generator().manual_seed(3)
    latent = torch.randn(16, 14, 512, generator=g).cuda()
    w = torch.randn(256, 512, generator=g).cuda()
    b = torch.randn(256, generator=g).cuda()
    (y,) = op.grouped_linear([(latent[:, 5], w, b, 0.0442, 1.0, False)])
    ref = latent[:, 5].double() @ w.double().t() * 0.0442 + b.double()
    assert _rel(y, ref) < 3e-5
    x = torch.randn(32, 8192, generator=g).cuda()
    w2 = torch.randn(512, 8192, generator=g).cuda()
    (y2,) = op.grouped_linear([(x, w2, None, 0.011, 1.0, False, 8)])
    assert _rel(y2, x.double() @ w2.double().t() * 0.011) < 3e-5


def test_grouped_linear_double_backward():
    from transeditor_b200 import op
    (x, w, b, alpha, bmul, act), = _layers([(8, 64, 24, True, True, 0.3, 0.5)])
    (y,) = op.grouped_linear([(x, w, b, alpha, bmul, act)])
    (gx,) = torch.autograd.grad(y.square().sum(), x, create_graph=True)
    gx.square().sum().backward()
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    r = op._linear_composite(xd, wd, bd, alpha, bmul, act, False)
    (gxd,) = torch.autograd.grad(r.square().sum(), xd, create_graph=True)
    gxd.square().sum().backward()
    assert _rel(w.grad, wd.grad) < 1e-4 and _rel(x.grad, xd.grad) < 1e-4


@pytest.mark.parametrize("batch,count,pn", [(16, 16, True), (8, 16, True), (3, 12, True), (16, 16, False), (33, 16, True)])
def test_mapping_columns(batch, count, pn):
    from transeditor_b200 import op
    g = torch.Generator().manual_seed(batch + count)
    code = torch.randn(batch, 512, 16, generator=g).cuda().requires_grad_(True)
    ws = [(torch.randn(512, 512, generator=g) * 100).cuda().requires_grad_(True) for _ in range(count)]
    bs = [torch.randn(512, generator=g).cuda().requires_grad_(True) for _ in range(count)]
    alpha, lr_mul = 0.01 / 512 ** 0.5, 0.01
    y = op.mapping_columns(code, ws, bs, alpha, lr_mul, pixel_norm=pn)
    cd = code.detach().double().requires_grad_(True)
    wd = [w.detach().double().requires_grad_(True) for w in ws]
    bd = [b.detach().double().requires_grad_(True) for b in bs]
    ref = op.mapping_columns_reference(cd, wd, bd, alpha, lr_mul, pn)
    assert _rel(y, ref) < 3e-5
    if count < 16:
        assert y[:, :, count:].abs().sum().item() == 0
    gy = torch.randn(y.shape, generator=g).cuda()
    (y * gy).sum().backward()
    (ref * gy.double()).sum().backward()
    assert _rel(code.grad, cd.grad) < 5e-5
    for a, r in zip(ws + bs, wd + bd):
        assert _rel(a.grad, r.grad) < 5e-5


def test_mapping_columns_double_backward_and_generator_route():
    from transeditor_b200 import op
    g = torch.Generator().manual_seed(1)
    code = torch.randn(4, 512, 16, generator=g).cuda().requires_grad_(True)
    ws = [(torch.randn(512, 512, generator=g) * 100).cuda().requires_grad_(True) for _ in range(16)]
    bs = [torch.zeros(512).cuda().requires_grad_(True) for _ in range(16)]
    y = op.mapping_columns(code, ws, bs, 0.01 / 512 ** 0.5, 0.01)
    (gc,) = torch.autograd.grad(y.square().sum(), code, create_graph=True)
    gc.square().sum().backward()
    assert ws[0].grad is not None and torch.isfinite(ws[0].grad).all()
