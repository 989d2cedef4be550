"""Split-operand (fp32-on-tensor-cores) mode of te_conv_tc / te_conv_wgrad_tc / te_split_bf16 /
te_pack_weights_tc against an f64 torch convolution of the SAME f32 operands.

Bars (stated per test): 2 planes = 3 products, each product exact to ~2^-16 -> max-abs error <= 6e-5 of the
output's scale (K-term random sums); 3 planes = 6 products -> the f32 accumulator's own rounding (<= 1e-5)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _randn(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.fixture(autouse=True)
def _planes():
    from transeditor_b200 import tc
    old = tc.get_split_planes()
    yield
    tc.set_split_planes(old)


def test_split_planes_reconstruct_the_input():
    from transeditor_b200 import tc
    x = _cl(_randn(3, 24, 9, 7, seed=1) * 37.5)
    s = _randn(3, 24, seed=2)
    for nseg, tol in ((1, 2 ** -8), (2, 2 ** -16), (3, 2 ** -22)):
        pl = tc.split_planes(x, nseg)
        assert pl.shape == (nseg, 3, 9, 7, 24) and pl.dtype == torch.bfloat16
        rec = pl.float().sum(0).permute(0, 3, 1, 2)
        assert ((rec - x).abs() <= tol * x.abs() + 1e-30).all()
        assert torch.equal(pl[0], x.permute(0, 2, 3, 1).to(torch.bfloat16))
    pl = tc.split_planes(x, 2, s)
    want = x * s[:, :, None, None]
    rec = pl.float().sum(0).permute(0, 3, 1, 2)
    assert ((rec - want).abs() <= 2 ** -16 * want.abs() + 1e-30).all()


def _ref(x, w, kind, k):
    x, w = x.double(), w.double()
    if kind == "s1":
        return F.conv2d(x, w, padding=k // 2)
    if kind == "down":
        return F.conv2d(x, w, stride=2)
    return F.conv_transpose2d(x, w.transpose(0, 1), stride=2)


@pytest.mark.parametrize("nseg,tol", [(2, 6e-5), (3, 2e-5)])
@pytest.mark.parametrize("kind,k,b,cin,cout,h", [
    ("s1", 3, 2, 64, 128, 32), ("s1", 1, 2, 72, 8, 16), ("down", 3, 2, 128, 64, 33), ("up", 3, 2, 64, 64, 16),
    ("s1", 3, 4, 128, 128, 64),      # enough tiles for the 2-CTA kernel
    ("s1", 3, 2, 256, 256, 32),      # BLOCK_N = 256
    ("s1", 3, 16, 512, 512, 4),      # several samples per tile
])
def test_conv_forward_and_gradients_match_f64(kind, k, b, cin, cout, h, nseg, tol):
    from transeditor_b200 import tc
    tc.set_split_planes(nseg)
    x = _cl(_randn(b, cin, h, h, seed=3)).requires_grad_(True)
    w = (_randn(cout, cin, k, k, seed=4) / (cin * k * k) ** 0.5).requires_grad_(True)
    if kind == "s1":
        y = tc.conv2d(x, w)
    elif kind == "down":
        y = tc.conv2d(x, w, stride=2)
    else:
        y = tc.conv_transpose2d(x, w)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = _ref(xr, wr, kind, k)
    assert y.dtype == torch.float32 and y.shape == yr.shape
    scale = yr.abs().max().item()
    assert (y.double() - yr).abs().max().item() < tol * scale
    g = _cl(_randn(*y.shape, seed=5))
    gx, gw = torch.autograd.grad(y, (x, w), g)
    gxr, gwr = torch.autograd.grad(yr, (xr, wr), g.double())
    assert (gx.double() - gxr).abs().max().item() < tol * gxr.abs().max().item()
    assert (gw.double() - gwr).abs().max().item() < tol * gwr.abs().max().item()


def test_per_sample_weights_bias_act_and_residual():
    from transeditor_b200 import tc
    b, c, h = 2, 64, 16
    x = _cl(_randn(b, c, h, h, seed=6))
    wb = _randn(b, c, c, 3, 3, seed=7) / (c * 9) ** 0.5
    tc.set_split_planes(2)
    y = tc.conv2d(x, wb)
    ref = torch.cat([F.conv2d(x[i:i + 1].double(), wb[i].double(), padding=1) for i in range(b)])
    assert (y.double() - ref).abs().max().item() < 6e-5 * ref.abs().max().item()
    w = _randn(c, c, 3, 3, seed=8) / (c * 9) ** 0.5
    bias = _randn(c, seed=9)
    res = _cl(_randn(b, c, h, h, seed=10))
    out = tc.conv2d_bias_act(x, w, bias)
    want = F.leaky_relu(F.conv2d(x.double(), w.double(), padding=1) + bias.double().view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert (out.double() - want).abs().max().item() < 6e-5 * want.abs().max().item()
    out = tc.conv2d_residual(x, w, res)
    want = F.conv2d(x.double(), w.double(), padding=1) + res.double()
    assert (out.double() - want).abs().max().item() < 6e-5 * want.abs().max().item()


def test_second_order_through_split_convs():
    """R1-style double backward: d/dw of |d y/d x|^2."""
    from transeditor_b200 import tc
    tc.set_split_planes(2)
    x = _cl(_randn(2, 32, 16, 16, seed=11)).requires_grad_(True)
    w = (_randn(32, 32, 3, 3, seed=12) / 17.0).requires_grad_(True)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)

    def penalty(xx, ww, conv):
        y = conv(xx, ww)
        (gx,) = torch.autograd.grad((y * y).sum(), xx, create_graph=True)
        return gx.pow(2).sum()

    p = penalty(x, w, lambda a, b_: tc.conv2d(a, b_))
    pr = penalty(xr, wr, lambda a, b_: F.conv2d(a, b_, padding=1))
    (gw,) = torch.autograd.grad(p, w)
    (gwr,) = torch.autograd.grad(pr, wr)
    assert abs(p.item() - pr.item()) < 1e-4 * abs(pr.item())
    assert (gw.double() - gwr).abs().max().item() < 2e-4 * gwr.abs().max().item()


def test_simt_and_split_generators_agree():
    """The exact-f32 SIMT engine and the split-operand tensor-core engine give the same image (< 1e-3)."""
    import model_spatial_query as M
    from transeditor_b200 import model as te_model
    torch.manual_seed(0)
    g = M.Generator(64, 512, 512, 10, channel_multiplier=2, n_trans=2, pixel_norm_op_dim=1).to(DEV)
    z, p = _randn(2, 512, 16, seed=13), _randn(2, 512, 16, seed=14)
    try:
        with torch.no_grad():
            te_model.set_precision("fp32_simt")
            a, _, _ = g(z, p)
            te_model.set_precision("fp32")
            b_, _, _ = g(z, p)
    finally:
        te_model.set_precision("fp32")
    assert (a - b_).abs().max().item() < 1e-3
