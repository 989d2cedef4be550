"""Size-independent properties checked at BASELINE.json's FULL sizes (256^2, batch 16), where the CPU oracle is too
slow to run: every operator on the path is linear in its data argument (and bilinear in data x weight), so forward
and backward kernels must satisfy the adjoint identity  <A x, y> = <x, A^T y>  exactly up to rounding; bias + LeakyReLU
is positively homogeneous; the generator is deterministic.  Small-size parity against the oracle lives in
test_gpu_ops.py / test_gpu_tc.py / test_gpu_model.py."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _randn(*shape, seed=0, dtype=torch.float32):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV, dtype=torch.float32).to(dtype)


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


def _noise(a, b):
    """Scale of the rounding noise of <a, b> when every a_i carries an independent relative error: sqrt(sum (a_i b_i)^2)."""
    return (a.double() * b.double()).square().sum().sqrt().item()


@pytest.mark.parametrize("dtype,cl", [(torch.float32, False), (torch.bfloat16, True)])
@pytest.mark.parametrize("geom", ["blur_after_upconv", "d_blur", "rgb_upsample", "skip_decimate"])
def test_upfirdn2d_adjoint_identity_at_flagship_sizes(dtype, cl, geom):
    """App. B.2 shapes: Blur [16,128,257,257]->256^2 (pad 1,1), D's Blur 256^2->257^2 (pad 2,2), the RGB skip
    Upsample [16,3,128^2]->256^2 (up 2, pad 2,1), the ResBlock skip's blur + decimation (down 2)."""
    from utils.op import upfirdn2d
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    fir = (k1[:, None] * k1[None, :])
    fir = (fir / fir.sum()).to(DEV)
    if geom == "blur_after_upconv":
        shape, kw = (16, 128, 257, 257), dict(kernel=fir * 4, pad=(1, 1))
    elif geom == "d_blur":
        shape, kw = (16, 128, 256, 256), dict(kernel=fir, pad=(2, 2))
    elif geom == "rgb_upsample":
        shape, kw = (16, 8 if cl else 3, 128, 128), dict(kernel=fir * 4, up=2, pad=(2, 1))
    else:
        shape, kw = (16, 128, 256, 256), dict(kernel=fir, down=2, pad=(1, 1))
    x = _randn(*shape, seed=1, dtype=dtype)
    if cl:
        x = _cl(x)
    x.requires_grad_(True)
    y = upfirdn2d(x, **kw)
    g = _randn(*y.shape, seed=2, dtype=dtype)
    if cl:
        g = _cl(g)
    (gx,) = torch.autograd.grad(y, x, g)
    lhs, rhs = _dot(y, g), _dot(x, gx)
    # each side carries one output rounding per element (2^-9 relative in bf16, 2^-24 plus the 16-tap f32 sum in f32):
    # allow ~10 sigma of that noise
    tol = (1.2e-2 if dtype == torch.bfloat16 else 2e-6) * (_noise(y, g) + _noise(gx, x))
    assert abs(lhs - rhs) < tol, (geom, lhs, rhs, tol)


@pytest.mark.parametrize("case", [(128, 128, 256, "s1"), (256, 256, 128, "s1"), (512, 512, 64, "s1"),
                                  (128, 256, 255, "down"), (256, 128, 128, "up")])
def test_conv_tc_adjoint_and_bilinear_identities_at_flagship_sizes(case):
    """<conv(x; W), g> = <x, dgrad(g; W)> = <W, wgrad(x, g)> for the three flagship layers (309 GFLOP each) and the
    strided / transposed geometries, batch 16, through TcConv's autograd (all three tcgen05 kernels)."""
    from transeditor_b200 import tc
    cin, cout, h, kind = case
    x = _cl(_randn(16, cin, h, h, seed=3, dtype=torch.bfloat16)).requires_grad_(True)
    w = (_randn(cout, cin, 3, 3, seed=4) / math.sqrt(cin * 9)).requires_grad_(True)
    y = tc.conv_transpose2d(x, w) if kind == "up" else tc.conv2d(x, w, stride=1 if kind == "s1" else 2)
    g = _cl(_randn(*y.shape, seed=5, dtype=torch.bfloat16))
    gx, gw = torch.autograd.grad(y, (x, w), g)
    lhs = _dot(y, g)
    # y and gx are rounded to bf16 once per element (~10 sigma allowed); gw is f32.  W enters the kernels rounded to
    # bf16: the bilinear identity is taken against the weights the kernel saw
    wb = w.detach().to(torch.bfloat16)
    assert abs(lhs - _dot(x, gx)) < 1.2e-2 * (_noise(y, g) + _noise(gx, x)), (case, lhs, _dot(x, gx))
    assert abs(lhs - _dot(wb, gw)) < 1.2e-2 * _noise(y, g) + 1e-4 * _noise(gw, wb), (case, lhs, _dot(wb, gw))


@pytest.mark.parametrize("dtype,cl", [(torch.float32, False), (torch.bfloat16, True)])
def test_fused_leaky_relu_homogeneity_and_bias_gradient_at_flagship_size(dtype, cl):
    """[16,128,256,256]: f(a x, a b) = a f(x, b) for a = 2^k (exact in floating point), and the bias gradient equals
    the per-channel sum of the input gradient (the one-pass backward kernel's two outputs agree)."""
    from utils.op import fused_leaky_relu
    x = _randn(16, 128, 256, 256, seed=6, dtype=dtype)
    if cl:
        x = _cl(x)
    b = _randn(128, seed=7, dtype=dtype)
    with torch.no_grad():
        y1 = fused_leaky_relu(x, b)
        y4 = fused_leaky_relu(x * 4, b * 4)
    assert torch.equal(y4, y1 * 4)
    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = fused_leaky_relu(xr, br)
    g = torch.ones_like(y)
    gx, gb = torch.autograd.grad(y, (xr, br), g)
    ref = gx.float().sum((0, 2, 3))
    assert (gb.float() - ref).abs().max().item() < (2e-2 if dtype == torch.bfloat16 else 1e-4) * ref.abs().max().item()


def test_attn_stack_directional_derivative_at_batch_16():
    """te_attn_stack_bwd against a central difference of te_attn_stack_fwd along a random direction (inputs and one
    weight matrix per block), batch 16, all 8 blocks."""
    from tests.test_gpu_attn_stack import LR, _blocks, _inputs, _to
    from transeditor_b200 import op
    bd = _to(_blocks(8, 528, seed=21), torch.float32, DEV, grad=True)
    x0, p0, p = [t.float().to(DEV).requires_grad_(True) for t in _inputs(16, 528, seed=22)]
    gy = _randn(16, 16, 512, seed=23)
    ws = [blk["w_m1"] for blk in bd]
    y = op.attn_stack(x0, p0, p, bd, LR)
    grads = torch.autograd.grad(y, [x0, p0, p] + ws, gy)
    dirs = [_randn(*t.shape, seed=30 + i) for i, t in enumerate([x0, p0, p] + ws)]
    eps = 1e-3

    def shifted(sign):
        with torch.no_grad():
            for t, d in zip([x0, p0, p] + ws, dirs):
                t.add_(d, alpha=sign * eps * (1.0 if t.dim() == 3 else 1.0 / LR))
            out = op.attn_stack(x0, p0, p, bd, LR)
            for t, d in zip([x0, p0, p] + ws, dirs):
                t.sub_(d, alpha=sign * eps * (1.0 if t.dim() == 3 else 1.0 / LR))
        return out

    # the weight directions are scaled by 1/lr_mul like the weights themselves (EqualLinear stores W / lr_mul)
    analytic = sum(_dot(g, d) * (1.0 if d.dim() == 3 else 1.0 / LR) for g, d in zip(grads, dirs))
    numeric = _dot(shifted(+1) - shifted(-1), gy) / (2 * eps)
    assert abs(analytic - numeric) < 2e-2 * max(abs(analytic), abs(numeric)), (analytic, numeric)


def test_generator_bf16_is_deterministic_at_batch_16():
    import model_spatial_query as M
    from transeditor_b200 import model as te_model
    te_model.set_precision("bf16")
    try:
        torch.manual_seed(0)
        g = M.Generator(256, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV).eval()
        z, p = _randn(16, 512, 16, seed=40), _randn(16, 512, 16, seed=41)
        with torch.no_grad():
            a = g(z, p, randomize_noise=False)[0]
            b = g(z, p, randomize_noise=False)[0]
        assert a.shape == (16, 3, 256, 256) and torch.isfinite(a).all()
        assert torch.equal(a, b)
    finally:
        te_model.set_precision("fp32")
