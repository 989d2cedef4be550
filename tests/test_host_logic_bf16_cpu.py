"""The bf16 tensor-core ROUTE's host logic on CPU: tests/emu.py emulates te_conv_tc / te_conv_wgrad_tc (and the other
entry points) from the header contract, so the tap-table geometry algebra of transeditor_b200/tc.py (adjoint modes,
polyphase transposed convolution, weight-gradient layouts, per-sample weights, epilogues), the bf16 branches of
model.py, the weight-repack cache and the trainer's phases are checked without a GPU — against torch convolutions
and against the reference's golden vectors.  The kernels themselves are checked on the GPU (test_gpu_tc.py, ...)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import te_oracle as O
from tests.conftest import load_golden, small

pytestmark = pytest.mark.usefixtures("cpu_emulation")


@pytest.fixture
def bf16_mode():
    from transeditor_b200 import model
    model.set_precision("bf16")
    yield
    model.set_precision("fp32")


def _bf(x):
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _cos(a, b):
    a, b = a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("kind,k,h", [("s1", 3, 8), ("s1", 1, 6), ("down", 3, 9), ("down", 1, 7), ("up", 3, 5), ("up", 1, 4)])
@pytest.mark.parametrize("per_sample", [False, True])
def test_tc_modes_forward_dgrad_wgrad_match_torch(kind, k, h, per_sample):
    from transeditor_b200 import tc
    b, cin, cout = 2, 16, 24
    x = _bf(_rand(b, cin, h, h, seed=1)).requires_grad_(True)
    wshape = (b, cout, cin, k, k) if per_sample else (cout, cin, k, k)
    w = _rand(*wshape, seed=2, scale=1 / math.sqrt(cin * k * k)).requires_grad_(True)
    y = tc.conv_transpose2d(x, w) if kind == "up" else tc.conv2d(x, w, stride=1 if kind == "s1" else 2)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).float().requires_grad_(True)

    def ref_one(xi, wi):
        if kind == "up":
            return F.conv_transpose2d(xi, wi.transpose(0, 1), stride=2)
        return F.conv2d(xi, wi, stride=1 if kind == "s1" else 2, padding=k // 2 if kind == "s1" else 0)

    ref = torch.cat([ref_one(xr[i:i + 1], wr[i]) for i in range(b)]) if per_sample else ref_one(xr, wr)
    assert y.shape == ref.shape
    g = _bf(_rand(*ref.shape, seed=3))
    gx, gw = torch.autograd.grad(y, (x, w), g)
    rx, rw = torch.autograd.grad(ref, (xr, wr), g.float())
    assert (y.float() - ref).abs().max().item() < 1.2e-2 * max(1.0, ref.abs().max().item())
    assert (gx.float() - rx).abs().max().item() < 1.2e-2 * max(1.0, rx.abs().max().item())
    assert (gw - rw).abs().max().item() < 1e-4 * max(1.0, rw.abs().max().item())


def test_tc_inference_geometries_and_epilogues():
    """down1 (stride 2, padding 1), bias + LeakyReLU(0.01), per-channel PReLU, residual."""
    from transeditor_b200 import tc
    x = _bf(_rand(2, 16, 10, 10, seed=4))
    w = _rand(24, 16, 3, 3, seed=5, scale=1 / 12)
    bias, slope = _rand(24, seed=6, scale=0.3), torch.rand(24, generator=torch.Generator().manual_seed(7))
    wf = w.to(torch.bfloat16).float()
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode("down1", 3), bias=bias, act=2)
    ref = F.leaky_relu(F.conv2d(x.float(), wf, bias, stride=2, padding=1), 0.01)
    assert (y.float() - ref).abs().max().item() < 1.2e-2
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode("s1", 3), act=3, slope=slope)
    ref = F.prelu(F.conv2d(x.float(), wf, padding=1), slope)
    assert (y.float() - ref).abs().max().item() < 1.2e-2
    res = _bf(_rand(2, 24, 10, 10, seed=8))
    y = tc.conv_raw(x, tc.pack_weight(w, False), tc.Mode("s1", 3), bias=bias, act=True, act_gain=0.5, residual=res)
    ref = F.leaky_relu(F.conv2d(x.float(), wf, bias, padding=1), 0.2) * 0.5 + res.float()
    assert (y.float() - ref).abs().max().item() < 2e-2


def _models(size, cm):
    import model_spatial_query as M
    t = 2 * int(np.log2(size)) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1)
    d = M.Discriminator(size, channel_multiplier=cm)
    g.load_state_dict(O.synthetic_state(O.generator_shapes(size, cm)), strict=True)
    d.load_state_dict(O.synthetic_state(O.discriminator_shapes(size, cm)), strict=True)
    return g, d


def test_bf16_route_forward_and_gradients_follow_reference(bf16_mode):
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    with torch.no_grad():
        img, _, _ = g(z, p)
        pred = d(torch.from_numpy(gold["real"])).numpy()
    std = gold["img"].std()
    err = np.abs(img.numpy() - gold["img"])
    assert err.mean() < 0.02 * std and err.max() < 0.15 * std, (err.mean(), err.max(), std)
    assert np.abs(pred - gold["d_real"]).max() < 0.05 * max(1.0, np.abs(gold["d_real"]).max())
    img, lat, _ = g(z, p, return_latents=True)
    loss = F.softplus(-d(img)).mean()
    loss.backward()
    assert abs(loss.item() - float(gold["g_loss"])) < 0.05 * max(1.0, abs(float(gold["g_loss"])))
    gp, dp = dict(g.named_parameters()), dict(d.named_parameters())
    for key, val in gold.items():
        if key.startswith("ggrad."):
            got = small(gp[key[6:]].grad)
        elif key.startswith("dgrad_from_g."):
            got = small(dp[key[13:]].grad)
        else:
            continue
        assert _cos(got, val) > 0.98, (key, _cos(got, val))


def test_bf16_route_regularisers_double_backward(bf16_mode):
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    real = torch.from_numpy(gold["real"]).requires_grad_(True)
    (gi,) = torch.autograd.grad(d(real).sum(), real, create_graph=True)
    r1 = gi.pow(2).reshape(gi.shape[0], -1).sum(1).mean()
    r1.backward()
    assert abs(r1.item() - float(gold["r1"])) < 0.1 * abs(float(gold["r1"]))
    dp, gp = dict(d.named_parameters()), dict(g.named_parameters())
    for key, val in gold.items():
        if key.startswith("r1grad."):
            assert _cos(small(dp[key[7:]].grad), val) > 0.95, key
    img, lat, _ = g(torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"]), return_latents=True)
    (gl,) = torch.autograd.grad((img * torch.from_numpy(gold["path_noise"])).sum(), lat, create_graph=True)
    pl = torch.sqrt(gl.pow(2).sum(2).mean(1))
    (pl - 0.5).pow(2).mean().backward()
    assert np.abs(pl.detach().numpy() / gold["path_lengths"] - 1).max() < 0.1
    for key, val in gold.items():
        if key.startswith("pathgrad."):
            assert _cos(small(gp[key[9:]].grad), val) > 0.95, key


def test_bf16_trainer_steps_with_the_weight_repack_cache(bf16_mode):
    """Two iterations (all four phases in the first): the cached bf16 weight copies must follow the fused Adam's
    raw-pointer updates — compare with a trainer whose cache is switched off."""
    from transeditor_b200 import tc
    from transeditor_b200.train_step import TrainConfig, Trainer
    real = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1
    runs = []
    for cached in (True, False):
        tr = Trainer(TrainConfig(size=32, batch=4), "cpu", seed=0)
        if not cached:
            tr._packs = tc.PackCache()  # nothing registered: every lookup misses
        for _ in range(2):
            tr.step(real)
        assert torch.isfinite(tr.g_flat.data).all() and torch.isfinite(tr.d_flat.data).all()
        runs.append((tr.g_flat.data.clone(), tr.d_flat.data.clone(), len(tr._packs._entries)))
    tc.set_pack_cache(None)
    assert runs[0][2] > 10 and runs[1][2] == 0
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
