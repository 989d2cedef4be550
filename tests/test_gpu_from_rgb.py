"""te_from_rgb_fwd / te_from_rgb_bwd (the discriminator's first layer) against float64 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(b, c, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    img = (torch.rand(b, 3, h, w, generator=g) * 2 - 1).cuda().requires_grad_(True)
    weight = torch.randn(c, 3, 1, 1, generator=g).cuda().requires_grad_(True)
    bias = (torch.randn(c, generator=g) * 0.3).cuda().requires_grad_(True)
    return img, weight, bias


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("shape", [(2, 128, 32, 32), (3, 32, 17, 9), (1, 256, 8, 8), (4, 8, 16, 16)])
def test_from_rgb_forward_backward(dtype, tol, shape):
    from transeditor_b200 import op
    b, c, h, w = shape
    img, weight, bias = _case(b, c, h, w)
    wscale = 1 / 3 ** 0.5
    out = op.from_rgb(img, weight, bias, wscale, dtype=dtype)
    assert out.dtype == dtype and out.is_contiguous(memory_format=torch.channels_last)
    ref_in = [t.detach().double().requires_grad_(True) for t in (img, weight, bias)]
    ref = op.from_rgb_reference(*ref_in, wscale, 2 ** 0.5)
    assert (out.double() - ref).abs().max().item() < tol * ref.abs().max().item()
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).cuda()
    gq = g.to(dtype)
    (out * gq).sum().backward()
    # the mask comes from the kernel's own (rounded) output: differentiate the reference with that mask
    mask = torch.where(out.double() > 0, 2 ** 0.5, 0.2 * 2 ** 0.5)
    gp = gq.double() * mask
    gw_ref = torch.einsum("bohw,bchw->oc", gp, img.detach().double()) * wscale
    gb_ref = gp.sum((0, 2, 3))
    gx_ref = torch.einsum("bohw,oc->bchw", gp, weight.detach().double().view(c, 3)) * wscale
    rel = lambda a, r: ((a.double() - r).abs().max() / r.abs().max()).item()  # noqa: E731
    assert rel(weight.grad.view(c, 3), gw_ref) < 1e-4
    assert rel(bias.grad, gb_ref) < 1e-4
    assert rel(img.grad, gx_ref) < 1e-5


def test_from_rgb_full_size_and_double_backward():
    from transeditor_b200 import op
    img, weight, bias = _case(16, 128, 256, 256, seed=2)
    out = op.from_rgb(img, weight, bias, 0.577)
    ref = op.from_rgb_reference(img[:2].detach(), weight.detach(), bias.detach(), 0.577, 2 ** 0.5)
    assert (out[:2].float() - ref).abs().max().item() < 0.05
    # linearity of the backward pass in g (same mask): bwd(2 g) == 2 bwd(g)
    g = torch.randn_like(out)
    (gi1,) = torch.autograd.grad(out, img, g, retain_graph=True)
    (gi2,) = torch.autograd.grad(out, img, 2 * g)
    assert torch.allclose(gi2, 2 * gi1, rtol=1e-5, atol=1e-6)
    # second order (the R1 route): d/dW of |d out / d img|^2 exists and matches the composite
    img2, w2, b2 = _case(2, 32, 8, 8, seed=3)
    o = op.from_rgb(img2, w2, b2, 0.5, dtype=torch.float32)
    (gi,) = torch.autograd.grad(o.sum(), img2, create_graph=True)
    gi.square().sum().backward()
    ref_in = [t.detach().double().requires_grad_(True) for t in (img2, w2, b2)]
    r = op.from_rgb_reference(*ref_in, 0.5, 2 ** 0.5)
    (gr,) = torch.autograd.grad(r.sum(), ref_in[0], create_graph=True)
    gr.square().sum().backward()
    # the second-order route runs on the split-operand tensor-core ops (~1e-5 relative)
    err = (w2.grad.double() - ref_in[1].grad).abs().max() / ref_in[1].grad.abs().max()
    assert err.item() < 1e-3
