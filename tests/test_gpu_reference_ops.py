"""te_upfirdn2d / te_fused_bias_act against the REFERENCE'S OWN CUDA kernels (utils/op/upfirdn2d_kernel.cu,
fused_bias_act_kernel.cu, JIT-built for sm_100a the way utils/op/fused_act.py:9-15 and upfirdn2d.py:8-14 build them),
and the Generator / Discriminator against the reference's classes on the same GPU.  Needs the unmodified reference
tree on the box (baseline/_ref, installed by tools/install_reference.py and shipped by gpurun); skipped without it.
Bar: f32, max-abs <= 1e-5 of the output scale per op (both sides are f32 kernels with different summation orders);
model outputs < 1e-3 (BASELINE.json)."""
import numpy as np
import pytest
import torch

from oracle import ref_gpu, te_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ref_ops():
    if ref_gpu.reference_root() is None:
        pytest.skip("no reference tree on this machine (baseline/_ref)")
    try:
        return ref_gpu.load_reference_ops()
    except Exception as e:  # toolchain trouble must not fail the product's suite
        pytest.skip("reference CUDA extensions did not build: %s" % str(e)[:200])


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(DEV)


@pytest.mark.parametrize("case", [
    # (n, c, h, w, up, down, pad): the live parameter sets of the path and their gradients
    (2, 8, 33, 33, 1, 1, (1, 1)), (2, 8, 32, 32, 1, 1, (2, 2)), (2, 3, 16, 16, 2, 1, (2, 1)), (2, 8, 32, 32, 1, 2, (1, 1)),
    (1, 4, 64, 48, 1, 1, (2, 1)), (4, 16, 65, 65, 1, 1, (1, 1)),
])
def test_upfirdn2d_matches_the_reference_kernel(ref_ops, case):
    from utils.op import upfirdn2d
    _, ref_up = ref_ops
    n, c, h, w, up, down, pad = case
    fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.]))
    fir = (fir / fir.sum() * up * up).to(DEV)
    outs = []
    for fn in (upfirdn2d, ref_up.upfirdn2d):
        x = _rand(n, c, h, w, seed=1).requires_grad_(True)
        y = fn(x, fir, up=up, down=down, pad=pad)
        gy = _rand(*y.shape, seed=2).requires_grad_(True)
        (gx,) = torch.autograd.grad(y, x, gy, create_graph=True)
        (gg,) = torch.autograd.grad(gx, gy, _rand(*gx.shape, seed=3))
        outs.append((y.detach(), gx.detach(), gg.detach()))
    for a, r in zip(*outs):
        assert a.shape == r.shape
        assert (a - r).abs().max().item() <= 1e-5 * max(1.0, r.abs().max().item())


def test_fused_leaky_relu_matches_the_reference_kernel(ref_ops):
    from utils.op import fused_leaky_relu
    ref_fa, _ = ref_ops
    outs = []
    for fn in (fused_leaky_relu, ref_fa.fused_leaky_relu):
        x = _rand(4, 16, 32, 32, seed=4).requires_grad_(True)
        b = _rand(16, seed=5).requires_grad_(True)
        y = fn(x, b)
        gy = _rand(*y.shape, seed=6).requires_grad_(True)
        gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
        (gg,) = torch.autograd.grad((gx * _rand(*gx.shape, seed=7)).sum() + (gb * _rand(16, seed=8)).sum(), gy)
        outs.append((y.detach(), gx.detach(), gb.detach(), gg.detach()))
    for a, r in zip(*outs):
        assert (a - r).abs().max().item() <= 1e-5 * max(1.0, r.abs().max().item())


def test_flagship_blur_equals_the_reference_kernel_and_is_faster(ref_ops):
    """[16, 128, 257, 257] -> 256^2 (BASELINE's upfirdn2d flagship) in the reference's own layout and dtype."""
    from utils.op import upfirdn2d
    _, ref_up = ref_ops
    fir = torch.outer(torch.tensor([1., 3., 3., 1.]), torch.tensor([1., 3., 3., 1.]))
    fir = (fir / fir.sum()).to(DEV)
    x = _rand(16, 128, 257, 257, seed=9)
    a, r = upfirdn2d(x, fir, pad=(1, 1)), ref_up.upfirdn2d(x, fir, pad=(1, 1))
    assert (a - r).abs().max().item() <= 1e-5 * max(1.0, r.abs().max().item())

    def ms(fn):
        for _ in range(2):
            fn(x, fir, pad=(1, 1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(x, fir, pad=(1, 1))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    ours, theirs = ms(upfirdn2d), ms(ref_up.upfirdn2d)
    print("upfirdn2d flagship f32 NCHW: ours %.3f ms, reference kernel %.3f ms" % (ours, theirs))
    assert ours < theirs


@pytest.mark.parametrize("size,batch", [(64, 2), (256, 1)])
def test_models_match_the_reference_classes_on_the_gpu(ref_ops, size, batch):
    import model_spatial_query as M
    ref = ref_gpu.load_reference_model()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    t = 2 * int(np.log2(size)) - 2
    sdg = O.synthetic_state(O.generator_shapes(size, 2))
    sdd = O.synthetic_state(O.discriminator_shapes(size, 2))
    outs = []
    for mod in (M, ref):
        g = mod.Generator(size, 512, 512, t, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV).eval()
        d = mod.Discriminator(size, channel_multiplier=2).to(DEV).eval()
        g.load_state_dict(sdg, strict=True)
        d.load_state_dict(sdd, strict=True)
        z, p = _rand(batch, 512, 16, seed=11), _rand(batch, 512, 16, seed=12)
        with torch.no_grad():
            img, lat, _ = g(z, p, return_latents=True)
            outs.append((img, lat, d(img)))
    for a, r in zip(*outs):
        assert (a - r).abs().max().item() < 1e-3


def test_full_baseline_size_against_the_reference_classes(ref_ops):
    """BASELINE configs[1] size (256^2, batch 16): forward, G-step gradients and D-step gradients of THIS repository's
    modules against the reference's own classes (its CUDA ops + cuDNN in true fp32) on the same GPU with the same
    weights and inputs.  fp32 parity mode (split-operand tensor-core convolutions): image / logits < 1e-3, gradients
    within 5e-3 of the tensor's max; bf16 mode: image mean-abs error < 2 % of the image std."""
    import model_spatial_query as M
    import torch.nn.functional as F
    from transeditor_b200 import model as te_model
    ref = ref_gpu.load_reference_model()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    size, batch = 256, 16
    sdg = O.synthetic_state(O.generator_shapes(size, 2))
    sdd = O.synthetic_state(O.discriminator_shapes(size, 2))
    z, p = _rand(batch, 512, 16, seed=21), _rand(batch, 512, 16, seed=22)
    real = _rand(batch, 3, size, size, seed=23).clamp(-1, 1)
    probe = ["convs.11.conv.weight", "convs.0.conv.weight", "conv1.conv.modulation.weight", "to_rgbs.5.conv.weight",
             "interact.3.mlp.0.weight", "spatial_mapping_network.5.weight", "convs.11.activate.bias"]
    dprobe = ["convs.0.0.weight", "convs.1.conv1.0.weight", "convs.6.skip.1.weight", "final_linear.0.weight", "convs.0.1.bias"]

    def run(mod):
        g = mod.Generator(size, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV)
        d = mod.Discriminator(size, channel_multiplier=2).to(DEV)
        g.load_state_dict(sdg, strict=True)
        d.load_state_dict(sdd, strict=True)
        for q in d.parameters():
            q.requires_grad_(False)
        img, _, _ = g(z, p)
        g_loss = F.softplus(-d(img)).mean()            # train_spatial_query.py:86-89
        g_loss.backward()
        gp = dict(g.named_parameters())
        g_grads = {k: gp[k].grad.detach().clone() for k in probe}
        for q in d.parameters():
            q.requires_grad_(True)
        fake = img.detach()
        d_loss = F.softplus(-d(real)).mean() + F.softplus(d(fake)).mean()   # :69-73
        d_loss.backward()
        dp = dict(d.named_parameters())
        d_grads = {k: dp[k].grad.detach().clone() for k in dprobe}
        return img.detach(), float(g_loss), float(d_loss), g_grads, d_grads

    te_model.set_precision("fp32")
    try:
        theirs = run(ref)
        ours = run(M)
        assert (ours[0] - theirs[0]).abs().max().item() < 1e-3
        assert abs(ours[1] - theirs[1]) < 1e-3 and abs(ours[2] - theirs[2]) < 1e-3
        for got, want in ((ours[3], theirs[3]), (ours[4], theirs[4])):
            for k in want:
                err = (got[k] - want[k]).abs().max().item() / max(want[k].abs().max().item(), 1e-20)
                assert err < 5e-3, (k, err)
        te_model.set_precision("bf16")
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.no_grad():
            g = M.Generator(size, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV).eval()
            g.load_state_dict(sdg, strict=True)
            img16 = g(z, p)[0].float()
        assert (img16 - theirs[0]).abs().mean().item() < 0.02 * theirs[0].std().item()
    finally:
        te_model.set_precision("fp32")
        torch.backends.cuda.matmul.allow_tf32 = False


def test_full_size_regularisers_against_the_reference_classes(ref_ops):
    """The two double-backward phases at BASELINE size against the reference's classes on the same GPU (fp32 parity mode):
    R1 penalty (train_spatial_query.py:76-83) at batch 16 and the path-length penalty (:92-105) at batch 8 — values and
    probed parameter gradients."""
    import model_spatial_query as M
    from transeditor_b200 import model as te_model
    ref = ref_gpu.load_reference_model()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    size = 256
    sdg = O.synthetic_state(O.generator_shapes(size, 2))
    sdd = O.synthetic_state(O.discriminator_shapes(size, 2))
    real = _rand(16, 3, size, size, seed=31).clamp(-1, 1)
    z, p = _rand(8, 512, 16, seed=32), _rand(8, 512, 16, seed=33)
    noise = _rand(8, 3, size, size, seed=34) / size
    dprobe = ["convs.0.0.weight", "convs.1.conv2.1.weight", "convs.3.conv1.0.weight", "final_conv.0.weight"]
    gprobe = ["convs.10.conv.weight", "convs.2.conv.weight", "convs.5.conv.modulation.weight", "to_rgb1.conv.weight"]

    def run(mod):
        d = mod.Discriminator(size, channel_multiplier=2).to(DEV)
        d.load_state_dict(sdd, strict=True)
        r = real.clone().requires_grad_(True)
        pred = d(r)
        r1 = O.d_r1_penalty(pred, r)
        (10 / 2 * r1 * 16 + 0 * pred[0]).backward()
        dp = dict(d.named_parameters())
        d_grads = {k: dp[k].grad.detach().clone() for k in dprobe}
        del d, pred, r
        g = mod.Generator(size, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV)
        g.load_state_dict(sdg, strict=True)
        img, lat, _ = g(z, p, return_latents=True)
        pl = O.g_path_lengths(img, lat, noise)
        pen = (pl - pl.mean().detach() * 0.5).pow(2).mean()
        (2 * 4 * pen + 0 * img[0, 0, 0, 0]).backward()
        gp = dict(g.named_parameters())
        g_grads = {k: gp[k].grad.detach().clone() for k in gprobe}
        return float(r1), d_grads, pl.detach().clone(), g_grads

    te_model.set_precision("fp32")
    theirs = run(ref)
    ours = run(M)
    assert abs(ours[0] - theirs[0]) < 2e-3 * max(abs(theirs[0]), 1e-6)
    assert (ours[2] - theirs[2]).abs().max().item() < 2e-3 * theirs[2].abs().max().item()
    for got, want in ((ours[1], theirs[1]), (ours[3], theirs[3])):
        for k in want:
            err = (got[k] - want[k]).abs().max().item() / max(want[k].abs().max().item(), 1e-20)
            assert err < 1e-2, (k, err)
