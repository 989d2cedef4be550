"""Tensor-core (tcgen05) convolution engine vs torch references computed in f32 from the same
bf16-rounded operands: every mode (stride-1, stride-2, transposed stride-2), data gradients, weight
gradients (MN-major tcgen05 kernel) and second order.  Tolerances = bf16 output rounding."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _bf(x):
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _ref(x, w, kind):
    if kind == "s1":
        return F.conv2d(x, w, padding=w.shape[2] // 2)
    if kind == "down":
        return F.conv2d(x, w, stride=2)
    return F.conv_transpose2d(x, w.transpose(0, 1), stride=2)


def _ours(x, w, kind):
    from transeditor_b200 import tc
    if kind == "up":
        return tc.conv_transpose2d(x, w)
    return tc.conv2d(x, w, stride=1 if kind == "s1" else 2)


def _close(a, r, what, rel=1.2e-2):
    a, r = a.float(), r.float()
    tol = rel * max(1.0, r.abs().max().item())
    err = (a - r).abs().max().item()
    assert err <= tol, "%s: err %.3e tol %.3e" % (what, err, tol)


FWD_CASES = [
    # b, cin, cout, h, w, k, kind
    (2, 64, 64, 16, 16, 3, "s1"), (3, 128, 64, 8, 8, 3, "s1"), (5, 64, 128, 4, 4, 3, "s1"),
    (2, 8, 64, 32, 32, 1, "s1"), (2, 64, 8, 32, 32, 1, "s1"), (1, 576, 512, 4, 4, 3, "s1"),
    (2, 64, 64, 17, 17, 3, "down"), (2, 128, 64, 33, 33, 3, "down"), (2, 64, 128, 15, 15, 1, "down"),
    (3, 64, 64, 9, 9, 3, "down"), (2, 64, 64, 8, 8, 3, "up"), (2, 128, 64, 16, 16, 3, "up"),
    (5, 64, 64, 4, 4, 3, "up"), (2, 64, 64, 8, 8, 1, "up"), (1, 64, 64, 64, 64, 3, "up"),
    (1, 64, 64, 129, 129, 3, "down"),
]


@pytest.mark.parametrize("case", FWD_CASES)
def test_tc_forward_modes(case):
    b, cin, cout, h, w_, k, kind = case
    x = _rand(b, cin, h, w_, seed=1)
    w = _rand(cout, cin, k, k, seed=2, scale=1.0 / math.sqrt(cin * k * k))
    xb, wb = _bf(x), w.to(torch.bfloat16).float()
    with torch.no_grad():
        y = _ours(xb, w, kind)
    ref = _ref(xb.float(), wb, kind)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    _close(y, ref, "fwd %s" % (case,))


GRAD_CASES = [
    (2, 64, 64, 16, 16, 3, "s1"), (2, 128, 64, 8, 8, 3, "s1"), (2, 64, 128, 32, 32, 1, "s1"),
    (2, 64, 64, 17, 17, 3, "down"), (2, 64, 128, 15, 15, 1, "down"), (2, 64, 64, 33, 33, 3, "down"),
    (2, 64, 64, 8, 8, 3, "up"), (2, 128, 64, 16, 16, 3, "up"), (4, 64, 64, 4, 4, 3, "up"),
    (2, 8, 64, 16, 16, 1, "s1"), (2, 64, 8, 16, 16, 1, "s1"),
]


@pytest.mark.parametrize("case", GRAD_CASES)
def test_tc_data_gradient(case):
    """dgrad runs the SAME kernel with transposed / flipped weight slices (mode.adjoint)."""
    b, cin, cout, h, w_, k, kind = case
    x = _rand(b, cin, h, w_, seed=3)
    w = _rand(cout, cin, k, k, seed=4, scale=1.0 / math.sqrt(cin * k * k))
    xb = _bf(x).requires_grad_(True)
    y = _ours(xb, w, kind)
    gy = _bf(_rand(*y.shape, seed=5))
    (gx,) = torch.autograd.grad(y, xb, gy)
    xr = xb.detach().float().requires_grad_(True)
    yr = _ref(xr, w.to(torch.bfloat16).float(), kind)
    (gxr,) = torch.autograd.grad(yr, xr, gy.float())
    assert gx.shape == gxr.shape
    _close(gx, gxr, "dgrad %s" % (case,))


@pytest.mark.parametrize("case", GRAD_CASES)
def test_tc_weight_gradient(case):
    b, cin, cout, h, w_, k, kind = case
    x = _rand(b, cin, h, w_, seed=6)
    w = _rand(cout, cin, k, k, seed=7, scale=1.0 / math.sqrt(cin * k * k)).requires_grad_(True)
    xb = _bf(x)
    y = _ours(xb, w, kind)
    gy = _bf(_rand(*y.shape, seed=8))
    (gw,) = torch.autograd.grad(y, w, gy)
    wr = w.detach().clone().requires_grad_(True)
    yr = _ref(xb.float(), wr, kind)
    (gwr,) = torch.autograd.grad(yr, wr, gy.float())
    assert gw.shape == gwr.shape and gw.dtype == torch.float32
    _close(gw, gwr, "wgrad %s" % (case,), rel=5e-3)


@pytest.mark.parametrize("case", [(2, 64, 64, 16, 16, 3, "s1"), (2, 64, 64, 17, 17, 3, "down"),
                                  (2, 64, 64, 8, 8, 3, "up"), (2, 64, 128, 15, 15, 1, "down")])
def test_tc_second_order(case):
    """R1 / path-length style double backward: grads of <gx, u> + <gw, v> w.r.t. gy, x and w."""
    b, cin, cout, h, w_, k, kind = case
    x = _rand(b, cin, h, w_, seed=9)
    w = _rand(cout, cin, k, k, seed=10, scale=1.0 / math.sqrt(cin * k * k))

    def run(fn, xin, win, cast):
        xx = xin.clone().requires_grad_(True)
        ww = win.clone().requires_grad_(True)
        y = fn(xx, ww, kind)
        gy = cast(_rand(*y.shape, seed=11)).requires_grad_(True)
        gx, gw = torch.autograd.grad(y, (xx, ww), gy, create_graph=True)
        u = cast(_rand(*gx.shape, seed=12))
        v = _rand(*gw.shape, seed=13)
        s = (gx.float() * u.float()).sum() + (gw * v).sum()
        return [t.detach().float() for t in torch.autograd.grad(s, (gy, xx, ww))]

    got = run(_ours, _bf(x), w, _bf)
    ref = run(_ref, _bf(x).float(), w.to(torch.bfloat16).float(), lambda t: _bf(t).float())
    for a, r, nm in zip(got, ref, ("h_gy", "h_x", "h_w")):
        _close(a, r, "%s %s" % (nm, case), rel=2e-2)


def test_tc_flagship_wgrad_and_dgrad_shapes():
    """BASELINE flagship layer 128->128 @256^2 B=16: gradient kernels at full size (linearity check)."""
    from transeditor_b200 import tc
    x = _bf(torch.randn(16, 128, 256, 256, device=DEV))
    w = torch.randn(128, 128, 3, 3, device=DEV) / 34
    gy = _bf(torch.randn(16, 128, 256, 256, device=DEV))
    mode = tc.Mode("s1", 3)
    gw = tc.wgrad_raw(gy, x, mode, (128, 128, 3, 3))
    gw2 = tc.wgrad_raw(gy[:8], x[:8], mode, (128, 128, 3, 3)) + tc.wgrad_raw(gy[8:], x[8:], mode, (128, 128, 3, 3))
    assert torch.isfinite(gw).all()
    assert (gw - gw2).abs().max().item() < 2e-3 * gw.abs().max().item()
    ref = torch.nn.grad.conv2d_weight(x[:1].float(), (128, 128, 3, 3), gy[:1].float(), padding=1)
    got = tc.wgrad_raw(gy[:1], x[:1], mode, (128, 128, 3, 3))
    _close(got, ref, "flagship wgrad sample 0", rel=5e-3)


@pytest.mark.parametrize("kind,h,k", [("s1", 16, 3), ("up", 16, 3), ("down", 33, 3), ("s1", 32, 1)])
def test_tc_per_sample_weights(kind, h, k):
    """Per-sample weights [B,O,I,K,K] (the reference's groups=batch formulation): forward, data gradient,
    per-sample weight gradient and second order against a per-sample loop of torch convolutions."""
    b, cin, cout = 3, 64, 64
    x = _bf(_rand(b, cin, h, h, seed=21))
    w = _rand(b, cout, cin, k, k, seed=22, scale=1.0 / math.sqrt(cin * k * k))

    def ref_fn(xx, ww, _kind):
        return torch.cat([_ref(xx[i:i + 1], ww[i], kind) for i in range(b)])

    def run(fn, xin, win, cast):
        xx = xin.clone().requires_grad_(True)
        ww = win.clone().requires_grad_(True)
        y = fn(xx, ww, kind)
        gy = cast(_rand(*y.shape, seed=23)).requires_grad_(True)
        gx, gw = torch.autograd.grad(y, (xx, ww), gy, create_graph=True)
        u = cast(_rand(*gx.shape, seed=24))
        v = _rand(*gw.shape, seed=25)
        s = (gx.float() * u.float()).sum() + (gw * v).sum()
        h_gy, h_x, h_w = torch.autograd.grad(s, (gy, xx, ww))
        return [t.detach().float() for t in (y, gx, gw, h_gy, h_x, h_w)]

    got = run(_ours, x, w, _bf)
    ref = run(ref_fn, x.float(), w.to(torch.bfloat16).float(), lambda t: _bf(t).float())
    for a, r, nm in zip(got, ref, ("y", "gx", "gw", "h_gy", "h_x", "h_w")):
        assert a.shape == r.shape, nm
        _close(a, r, "%s per-sample %s" % (nm, kind), rel=2e-2)


@pytest.mark.parametrize("case", [
    # large enough (>= 74 tile pairs) to take the 2-CTA (cta_group::2) kernel
    (16, 128, 128, 64, 64, 3, "s1"), (16, 256, 256, 64, 64, 3, "s1"), (11, 128, 128, 40, 40, 3, "s1"),
    (16, 128, 256, 65, 65, 3, "down"), (16, 256, 128, 32, 32, 3, "up"), (16, 64, 128, 64, 64, 1, "s1"),
])
def test_tc_two_cta_kernel(case):
    b, cin, cout, h, w_, k, kind = case
    x = _bf(_rand(b, cin, h, w_, seed=31))
    w = _rand(cout, cin, k, k, seed=32, scale=1.0 / math.sqrt(cin * k * k))
    xr = x.clone().requires_grad_(True)
    y = _ours(xr, w, kind)
    ref_in = x.float().requires_grad_(True)
    ref = _ref(ref_in, w.to(torch.bfloat16).float(), kind)
    assert y.shape == ref.shape
    _close(y, ref, "2cta fwd %s" % (case,))
    gy = _bf(_rand(*y.shape, seed=33))
    (gx,) = torch.autograd.grad(y, xr, gy)
    (gxr,) = torch.autograd.grad(ref, ref_in, gy.float())
    _close(gx, gxr, "2cta dgrad %s" % (case,))


def test_tc_two_cta_per_sample_and_epilogue():
    from transeditor_b200 import tc
    b, c, h, k = 16, 128, 64, 3
    x = _bf(_rand(b, c, h, h, seed=41))
    w = _rand(b, c, c, k, k, seed=42, scale=1.0 / math.sqrt(c * k * k))
    bias = _rand(c, seed=43)
    osc = _rand(b, c, seed=44).abs() + 0.5
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode("s1", k), out_scale=osc, bias=bias, act=True)
    wb = w.to(torch.bfloat16).float()
    ref = torch.cat([F.conv2d(x[i:i + 1].float(), wb[i], padding=1) for i in range(b)])
    ref = F.leaky_relu(ref * osc[:, :, None, None] + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    _close(y, ref, "2cta per-sample + epilogue")


@pytest.mark.parametrize("case", [(2, 64, 64, 16, 16, 3, "s1"), (16, 128, 256, 64, 64, 3, "s1"),
                                  (2, 64, 128, 33, 33, 3, "down"), (16, 256, 256, 63, 63, 1, "down"),
                                  (2, 64, 64, 8, 8, 3, "up")])
def test_tc_weight_scale_folded_into_repack(case):
    """conv(x, W * wscale) with the scale applied while the weight is repacked, and wscale * wgrad from the
    weight-gradient kernel (alpha): values and both gradients against torch."""
    from transeditor_b200 import tc
    b, cin, cout, h, w_, k, kind = case
    wscale = 1.0 / math.sqrt(cin * k * k)
    x = _bf(_rand(b, cin, h, w_, seed=51)).requires_grad_(True)
    w = _rand(cout, cin, k, k, seed=52).requires_grad_(True)
    if kind == "up":
        y = tc.conv_transpose2d(x, w, wscale=wscale)
    else:
        y = tc.conv2d(x, w, stride=1 if kind == "s1" else 2, wscale=wscale)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    ref = _ref(xr, (wr * wscale).to(torch.bfloat16).float(), kind)
    _close(y, ref, "wscale fwd %s" % (case,))
    gy = _bf(_rand(*y.shape, seed=53))
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    gxr, gwr = torch.autograd.grad(ref, (xr, wr), gy.float())
    _close(gx, gxr, "wscale dgrad %s" % (case,))
    _close(gw, gwr, "wscale wgrad %s" % (case,), rel=2e-2)


@pytest.mark.parametrize("case", [(2, 64, 64, 16, 16, 3, 1), (16, 128, 256, 64, 64, 1, 1), (4, 64, 128, 33, 33, 3, 2),
                                  (16, 256, 512, 63, 63, 1, 2), (3, 64, 72, 9, 9, 3, 1)])
def test_tc_conv_residual_epilogue(case):
    """y = conv(x, W * wscale) + residual with the sum in the epilogue (1-CTA and 2-CTA kernels); the residual's
    gradient is the incoming gradient itself."""
    from transeditor_b200 import tc
    b, cin, cout, h, w_, k, stride = case
    wscale = 0.7 / math.sqrt(cin * k * k)
    x = _bf(_rand(b, cin, h, w_, seed=61)).requires_grad_(True)
    w = _rand(cout, cin, k, k, seed=62).requires_grad_(True)
    ho = h if stride == 1 else (h - k) // 2 + 1
    wo = w_ if stride == 1 else (w_ - k) // 2 + 1
    res = _bf(_rand(b, cout, ho, wo, seed=63)).requires_grad_(True)
    y = tc.conv2d_residual(x, w, res, stride=stride, wscale=wscale)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    wq = (wr * wscale).to(torch.bfloat16).float()
    ref = (F.conv2d(xr, wq, padding=k // 2) if stride == 1 else F.conv2d(xr, wq, stride=2)) + res.detach().float()
    _close(y, ref, "residual fwd %s" % (case,))
    gy = _bf(_rand(*y.shape, seed=64))
    gx, gw, gr = torch.autograd.grad(y, (x, w, res), gy)
    gxr, gwr = torch.autograd.grad(ref, (xr, wr), gy.float())
    _close(gx, gxr, "residual dgrad %s" % (case,))
    _close(gw, gwr, "residual wgrad %s" % (case,), rel=2e-2)
    assert torch.equal(gr, gy)


@pytest.mark.parametrize("gain", [1.0, 2 ** -0.5 * 2 ** 0.5, 0.5])
def test_tc_bias_act_gain(gain):
    from transeditor_b200 import tc
    b, cin, cout, h, k = 4, 64, 128, 32, 3
    wscale = 1.0 / math.sqrt(cin * k * k)
    x = _bf(_rand(b, cin, h, h, seed=71)).requires_grad_(True)
    w = _rand(cout, cin, k, k, seed=72).requires_grad_(True)
    bias = _rand(cout, seed=73).requires_grad_(True)
    y = tc.conv2d_bias_act(x, w, bias, wscale=wscale, gain=gain)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    br = bias.detach().clone().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xr, (wr * wscale).to(torch.bfloat16).float(), padding=1) + br.view(1, -1, 1, 1), 0.2) * gain
    _close(y, ref, "gain fwd")
    gy = _bf(_rand(*y.shape, seed=74))
    gx, gw, gb = torch.autograd.grad(y, (x, w, bias), gy)
    gxr, gwr, gbr = torch.autograd.grad(ref, (xr, wr, br), gy.float())
    _close(gx, gxr, "gain dgrad")
    _close(gw, gwr, "gain wgrad", rel=2e-2)
    _close(gb, gbr, "gain bias grad", rel=2e-2)


def test_tc_carry_sums_fanout_gradient_in_dgrad_epilogue():
    """x feeds conv+act AND a second branch: with the carry form the second branch's gradient arrives as the
    dgrad kernel's residual; result and second order (R1-style) equal the plain two-consumer graph."""
    from transeditor_b200 import tc
    b, c, h, k = 4, 64, 32, 3
    wscale = 1.0 / math.sqrt(c * k * k)
    x0 = _bf(_rand(b, c, h, h, seed=81))
    w = _rand(c, c, k, k, seed=82).requires_grad_(True)
    bias = _rand(c, seed=83).requires_grad_(True)
    m2 = _bf(_rand(b, c, h, h, seed=84))
    gy = _bf(_rand(b, c, h, h, seed=85))

    def run(carry):
        x = x0.clone().requires_grad_(True)
        if carry:
            y, xa = tc.conv2d_bias_act_carry(x, w, bias, wscale=wscale)
        else:
            y, xa = tc.conv2d_bias_act(x, w, bias, wscale=wscale), x
        z = y + (xa * m2) * xa            # the second branch, non-linear in x so that second order is exercised
        (gx,) = torch.autograd.grad(z, x, gy, create_graph=True)
        pen = gx.float().square().mean()
        gw, gb = torch.autograd.grad(pen, (w, bias), allow_unused=True)
        return z, gx, gw, gb

    z0, gx0, gw0, gb0 = run(False)
    z1, gx1, gw1, gb1 = run(True)
    assert torch.equal(z0, z1)
    _close(gx1, gx0, "carry dgrad", rel=2e-2)
    _close(gw1, gw0, "carry second-order weight grad", rel=3e-2)
    if gb0 is not None:
        _close(gb1, gb0, "carry second-order bias grad", rel=3e-2)


@pytest.mark.parametrize("case", [(32, 512, 512, 16), (4, 512, 512, 2), (3, 64, 128, 64), (32, 512, 512, 4), (2, 128, 64, 33)])
def test_tc_padded_stride2_with_leaky001_epilogue(case):
    """Geometry `down1` (3x3, stride 2, padding 1) with bias + LeakyReLU(0.01) in the epilogue: the pSp heads'
    layer (pSp/models/encoders/psp_encoders_new.py:19-26)."""
    from transeditor_b200 import tc
    b, cin, cout, h = case
    x = _bf(_rand(b, cin, h, h, seed=1))
    w = _rand(cout, cin, 3, 3, seed=2, scale=1 / math.sqrt(cin * 9))
    bias = _rand(cout, seed=3, scale=0.3)
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode("down1", 3), bias=bias, act=2)
    ref = F.leaky_relu(F.conv2d(x.float(), w.to(torch.bfloat16).float(), bias, stride=2, padding=1), 0.01)
    assert y.shape == ref.shape
    _close(y, ref, "down1 %s" % (case,))


def test_pack_weights_table_matches_per_call_pack():
    """te_pack_weights_tc (all weights of a model, both orientations, one launch) == the per-call repack."""
    from transeditor_b200 import lib, tc
    shapes = [(512, 512, 3), (128, 3, 1), (3, 128, 1), (520, 512, 3), (512, 513, 3), (64, 128, 3), (256, 128, 1), (40, 72, 3)]
    tasks, want = [], []
    for n, (o, i, k) in enumerate(shapes):
        w = _rand(o, i, k, k, seed=n)
        scale = 1.0 / math.sqrt(i * k * k) if n % 2 else 1.0
        po, pi = (o + 7) // 8 * 8, (i + 7) // 8 * 8
        dn = torch.zeros(k * k, po, pi, dtype=torch.bfloat16, device=DEV) if n != 2 else None
        dt = torch.zeros(k * k, pi, po, dtype=torch.bfloat16, device=DEV) if n != 3 else None
        tasks.append((w, dn, dt, scale))
        want.append((tc.pack_weight(w, False, scale), tc.pack_weight(w, True, scale)))
    lib.pack_weights_tc(tasks * 10)   # 80 tasks: two launches
    for (w, dn, dt, scale), (rn, rt) in zip(tasks, want):
        if dn is not None:
            assert torch.equal(dn, rn), tuple(w.shape)
        if dt is not None:
            assert torch.equal(dt, rt), tuple(w.shape)


def test_pack_cache_serves_registered_weights_and_tracks_changes():
    from transeditor_b200 import tc
    w = torch.nn.Parameter(_rand(64, 32, 3, 3, seed=1))
    other = _rand(64, 32, 3, 3, seed=2)
    cache = tc.PackCache()
    cache.register([w])
    tc.set_pack_cache(cache)
    try:
        a = tc.pack_weight(w, False, 0.5)
        assert tc.pack_weight(w, False, 0.5) is a                      # served from the cache
        assert tc.pack_weight(other, False, 0.5) is not a              # unregistered: packed per call
        ref = a.clone()
        with torch.no_grad():
            w.copy_(other)                                             # torch-side change: version bump -> repacked
        b = tc.pack_weight(w, False, 0.5)
        assert b is a and not torch.equal(b, ref)
        tc.set_pack_cache(None)
        assert torch.equal(b, tc.pack_weight(other, False, 0.5))
        tc.set_pack_cache(cache)
        w.data.mul_(2.0)                                               # raw change (what the fused Adam does): explicit refresh
        cache.refresh()
        tc.set_pack_cache(None)
        assert torch.equal(a, tc.pack_weight(w.detach(), False, 0.5))
        t = cache.lookup(w, True, 0.5)
        assert torch.equal(t, tc.pack_weight(w.detach(), True, 0.5))
    finally:
        tc.set_pack_cache(None)


@pytest.mark.parametrize("case", [(4, 64, 64, 32, "s1", 3), (2, 128, 256, 16, "down1", 3), (3, 64, 128, 16, "down", 1), (2, 512, 512, 8, "s1", 3)])
def test_tc_prelu_epilogue(case):
    """act=3: per-channel PReLU slopes in the epilogue (the pSp trunk's conv + PReLU, helpers.py:87-91)."""
    from transeditor_b200 import tc
    b, cin, cout, h, kind, k = case
    x = _bf(_rand(b, cin, h, h, seed=1))
    w = _rand(cout, cin, k, k, seed=2, scale=1 / math.sqrt(cin * k * k))
    slope = torch.rand(cout, generator=torch.Generator().manual_seed(3)).to(DEV) * 0.5
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode(kind, k), act=3, slope=slope)
    stride, pad = {"s1": (1, k // 2), "down1": (2, 1), "down": (2, 0)}[kind]
    ref = F.prelu(F.conv2d(x.float(), w.to(torch.bfloat16).float(), stride=stride, padding=pad), slope)
    assert y.shape == ref.shape
    _close(y, ref, "prelu %s" % (case,))
