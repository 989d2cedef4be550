"""Edge cases of the operator API: empty tensors, non-contiguous inputs (the reference's native side
calls .contiguous(), fused_bias_act_kernel.cu:58-60, upfirdn2d_kernel.cu:149-150), 1x1 planes, the largest FIR the
ABI takes, odd batch sizes through the whole generator / discriminator."""
import pytest
import torch

from oracle import ops_cpu

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("shape", [(0, 8, 4, 4), (2, 8, 0, 4), (0, 16)])
def test_fused_leaky_relu_empty(shape):
    from utils.op import fused_leaky_relu
    x = torch.zeros(shape, device=DEV, requires_grad=True)
    b = torch.zeros(shape[1], device=DEV, requires_grad=True)
    y = fused_leaky_relu(x, b)
    assert y.shape == x.shape
    gx, gb = torch.autograd.grad(y.sum(), (x, b), allow_unused=True)
    assert gx is None or gx.shape == x.shape
    assert gb is None or (gb == 0).all()


def test_upfirdn2d_empty_batch_and_single_pixel():
    from utils.op import upfirdn2d
    k = torch.ones(4, 4, device=DEV) / 16
    y = upfirdn2d(torch.zeros(0, 3, 8, 8, device=DEV), k, up=2, pad=(2, 1))
    assert y.shape == (0, 3, 16, 16)
    x = _rand(2, 3, 1, 1, seed=1)
    y = upfirdn2d(x.to(DEV), k, up=2, pad=(2, 1))
    ref = ops_cpu.upfirdn2d(x.double(), k.cpu().double(), up=2, pad=(2, 1))
    assert y.shape == ref.shape == (2, 3, 2, 2)
    assert (y.cpu().double() - ref).abs().max().item() < 1e-6


def test_non_contiguous_inputs_match_contiguous():
    from utils.op import fused_leaky_relu, upfirdn2d
    base = _rand(2, 12, 10, 6, seed=2).to(DEV)
    xt = base.transpose(2, 3)                       # [2, 6... no: [2, 12, 6, 10] view with swapped strides
    b = _rand(12, seed=3).to(DEV)
    assert not xt.is_contiguous()
    assert torch.equal(fused_leaky_relu(xt, b), fused_leaky_relu(xt.contiguous(), b))
    k = torch.tensor([[1.0, 2.0], [3.0, 4.0]], device=DEV)
    assert torch.equal(upfirdn2d(xt, k, pad=(1, 0)), upfirdn2d(xt.contiguous(), k, pad=(1, 0)))
    sl = base[:, ::2]                               # channel-strided slice
    assert torch.equal(fused_leaky_relu(sl, b[::2]), fused_leaky_relu(sl.contiguous(), b[::2].contiguous()))


def test_upfirdn2d_largest_fir_and_too_large():
    from utils.op import upfirdn2d
    x = _rand(1, 2, 20, 20, seed=4)
    k = _rand(16, 16, seed=5)
    y = upfirdn2d(x.to(DEV), k.to(DEV), pad=(8, 7))
    ref = ops_cpu.upfirdn2d(x.double(), k.double(), pad=(8, 7))
    assert (y.cpu().double() - ref).abs().max().item() < 1e-4 * ref.abs().max().item()
    with pytest.raises(RuntimeError):
        upfirdn2d(x.to(DEV), _rand(17, 17, seed=6).to(DEV), pad=(8, 8))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("batch", [1, 3, 5])
def test_generator_and_discriminator_odd_batches(precision, batch):
    import model_spatial_query as M
    from transeditor_b200 import model as te_model
    te_model.set_precision(precision)
    try:
        torch.manual_seed(0)
        g = M.Generator(32, 512, 512, 8, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV)
        d = M.Discriminator(32, channel_multiplier=2).to(DEV)
        z, p = _rand(batch, 512, 16, seed=7).to(DEV), _rand(batch, 512, 16, seed=8).to(DEV)
        img, lat, _ = g(z, p, return_latents=True)
        assert img.shape == (batch, 3, 32, 32) and lat.shape == (batch, 8, 512)
        pred = d(img)      # minibatch-stddev group = min(batch, 4) must divide the batch: 1, 3 fine; 5 -> group 4 fails
        assert pred.shape == (batch, 1)
        pred.sum().backward()
        assert all(torch.isfinite(q.grad).all() for q in g.parameters() if q.grad is not None)
    except RuntimeError as e:
        # the reference fails the same way for batch % min(batch, 4) != 0 (model_spatial_query.py:845-847: view())
        assert batch == 5 and ("shape" in str(e) or "view" in str(e) or "invalid" in str(e)), e
    finally:
        te_model.set_precision("fp32")
