"""End-to-end parity of Generator / Discriminator / train step on the B200 with the reference's
golden vectors and the CPU oracle.  Bar: per-pixel max-abs < 1e-3 at fp32 (BASELINE.json)."""
import numpy as np
import pytest
import torch

from oracle import te_oracle as O
from tests.conftest import load_golden, small

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _models(size, cm, inject_noise=False):
    import model_spatial_query as M
    t = 2 * int(np.log2(size)) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1,
                    layer_noise_injection=inject_noise)
    d = M.Discriminator(size, channel_multiplier=cm)
    g.load_state_dict(O.synthetic_state(O.generator_shapes(size, cm)), strict=True)
    d.load_state_dict(O.synthetic_state(O.discriminator_shapes(size, cm)), strict=True)
    return g.to(DEV), d.to(DEV)


def _t(a):
    return torch.from_numpy(a).to(DEV)


@pytest.mark.parametrize("name", ["gd32_b4", "gd64_b2", "gd256_b1"])
def test_forward_parity_with_reference_golden(name):
    gold = load_golden(name)
    g, d = _models(int(gold["size"]), int(gold["cm"]))
    z, p = _t(gold["z"]), _t(gold["p"])
    with torch.no_grad():
        img, lat, _ = g(z, p, return_latents=True)
        assert np.abs(img.cpu().numpy() - gold["img"]).max() < 1e-3
        assert np.abs(lat.cpu().numpy() - gold["latent"]).max() < 1e-3
        zp, pp = g(z, p, return_mapped_codes=True)
        assert np.abs(zp.cpu().numpy() - gold["z_plus"]).max() < 1e-4
        assert np.abs(pp.cpu().numpy() - gold["p_plus"]).max() < 1e-4
        img2, _, _ = g(_t(gold["z_plus"]), _t(gold["p_plus"]), use_spatial_mapping=False, use_style_mapping=False)
        assert np.abs(img2.cpu().numpy() - gold["img_from_plus"]).max() < 1e-3
        assert np.abs(d(_t(gold["img"])).cpu().numpy() - gold["d_fake"]).max() < 1e-3
        assert np.abs(d(_t(gold["real"])).cpu().numpy() - gold["d_real"]).max() < 1e-3
    # the differentiable (unfused) route produces the same image
    zz = z.clone().requires_grad_(True)
    img3, _, _ = g(zz, p)
    assert np.abs(img3.detach().cpu().numpy() - gold["img"]).max() < 1e-3


def test_generator_step_gradients():
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    img, lat, _ = g(_t(gold["z"]), _t(gold["p"]), return_latents=True)
    loss = torch.nn.functional.softplus(-d(img)).mean()
    loss.backward()
    assert abs(loss.item() - float(gold["g_loss"])) < 1e-4
    gp, dp = dict(g.named_parameters()), dict(d.named_parameters())
    n = 0
    for key, val in gold.items():
        if key.startswith("ggrad."):
            got = small(gp[key[6:]].grad)
        elif key.startswith("dgrad_from_g."):
            got = small(dp[key[13:]].grad)
        else:
            continue
        n += 1
        assert np.abs(got - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key
    assert n >= 10


def test_r1_and_path_length_double_backward():
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    dp, gp = dict(d.named_parameters()), dict(g.named_parameters())
    real = _t(gold["real"]).requires_grad_(True)
    pred = d(real)
    (gi,) = torch.autograd.grad(pred.sum(), real, create_graph=True)
    r1 = gi.pow(2).reshape(gi.shape[0], -1).sum(1).mean()
    r1.backward()
    assert abs(r1.item() - float(gold["r1"])) < 1e-3 * max(1.0, abs(float(gold["r1"])))
    assert np.abs(gi.detach().cpu().numpy() - gold["r1_grad_img"]).max() < 1e-3 * max(1.0, np.abs(gold["r1_grad_img"]).max())
    for key, val in gold.items():
        if key.startswith("r1grad."):
            assert np.abs(small(dp[key[7:]].grad) - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key
    img, lat, _ = g(_t(gold["z"]), _t(gold["p"]), return_latents=True)
    (gl,) = torch.autograd.grad((img * _t(gold["path_noise"])).sum(), lat, create_graph=True)
    pl = torch.sqrt(gl.pow(2).sum(2).mean(1))
    (pl - 0.5).pow(2).mean().backward()
    assert np.abs(pl.detach().cpu().numpy() - gold["path_lengths"]).max() < 1e-3 * max(1.0, np.abs(gold["path_lengths"]).max())
    for key, val in gold.items():
        if key.startswith("pathgrad."):
            assert np.abs(small(gp[key[9:]].grad) - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key


def test_noise_injection_routes():
    gold = load_golden("g32_noise")
    g, _ = _models(32, 2, inject_noise=True)
    z, p = _t(gold["z"]), _t(gold["p"])
    noise = [_t(gold[f"noise_{i}"]) for i in range(7)]
    with torch.no_grad():
        img, _, _ = g(z, p, noise=noise)
        assert np.abs(img.cpu().numpy() - gold["img"]).max() < 1e-3
        img_b, _, _ = g(z, p, randomize_noise=False)
        assert np.abs(img_b.cpu().numpy() - gold["img_buffer_noise"]).max() < 1e-3
    img_g, _, _ = g(z.clone().requires_grad_(True), p, noise=noise)
    assert np.abs(img_g.detach().cpu().numpy() - gold["img"]).max() < 1e-3
    assert len(g.make_noise()) == 7


def test_full_size_batch16_properties():
    """BASELINE cfg-T size (256^2, B=16): finite output, batch independence of G (sample i of a batch
    equals the same latent run alone), D minibatch-stddev couples groups of 4 only."""
    import model_spatial_query as M
    torch.manual_seed(0)
    g = M.Generator(256, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV)
    d = M.Discriminator(256).to(DEV)
    z, p = torch.randn(16, 512, 16, device=DEV), torch.randn(16, 512, 16, device=DEV)
    with torch.no_grad():
        img, _, _ = g(z, p)
        assert img.shape == (16, 3, 256, 256) and torch.isfinite(img).all()
        one, _, _ = g(z[5:6], p[5:6])
        assert (one - img[5:6]).abs().max().item() < 1e-3
        pred = d(img)
        assert pred.shape == (16, 1) and torch.isfinite(pred).all()
        # stddev groups: samples {i, i+4, i+8, i+12} form one statistic (view(group, -1, ...), :846)
        img2 = img.clone()
        img2[1] = img2[1] * 0.5
        pred2 = d(img2)
        changed = (pred2 - pred).abs().squeeze(1) > 1e-6
        assert changed[1] and changed[5] and changed[9] and changed[13]
        assert not changed[0] and not changed[2] and not changed[3]


def test_train_step_matches_cpu_oracle_update():
    """One D step and one G step of the flat-buffer engine == the oracle's Adam update on the CPU."""
    from oracle.train_cpu import CpuTrainer
    from transeditor_b200.train_step import TrainConfig, Trainer
    cfg = TrainConfig(size=32, batch=4)
    tr = Trainer(cfg, DEV, seed=0)
    ref = CpuTrainer(size=32, batch=4)
    tr.generator.load_state_dict({k: v.detach() for k, v in ref.g.items()}, strict=True)
    tr.discriminator.load_state_dict({k: v.detach() for k, v in ref.d.items()}, strict=True)
    assert tr.g_flat.data.data_ptr() <= dict(tr.generator.named_parameters())["conv1.conv.weight"].data_ptr()
    gen = torch.Generator().manual_seed(99)
    lat = [torch.randn(4, 512, 16, generator=gen) for _ in range(4)]
    real = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
    it = iter(lat)
    tr._latents = lambda n: (next(it).to(DEV), next(it).to(DEV))
    tr.d_step(real.to(DEV))
    tr.g_step()
    # oracle side, same latents
    it2 = iter(lat)
    def fake(n, with_latent=False):
        z, p = next(it2), next(it2)
        img, l = O.generator_forward(ref.g, z, p, 32, 8)
        return (img, l) if with_latent else img
    ref._fake = fake
    ref.it = 1  # no lazy regularisers
    ref.step(real)
    dp = dict(tr.discriminator.named_parameters())
    gp = dict(tr.generator.named_parameters())
    def close(a, r, key):
        # Adam's first step is lr * g / (|g| + eps): elements whose gradient is ~eps are ill-conditioned,
        # so require agreement on all but a vanishing fraction and a tiny mean deviation.
        diff = (a - r).abs()
        assert (diff > 2e-4).float().mean().item() < 1e-3, key
        assert diff.mean().item() < 2e-5, key

    for k in ("final_linear.1.weight", "convs.0.1.bias", "convs.1.conv2.1.weight", "final_conv.0.weight"):
        close(dp[k].detach().cpu(), ref.d[k].detach(), k)
    for k in ("conv1.conv.weight", "to_rgb1.bias", "adjust_style.weight", "convs.1.activate.bias",
              "interact.3.mlp.0.weight", "style_mapping_network.5.weight"):
        close(gp[k].detach().cpu(), ref.g[k].detach(), k)
    # gradients were gathered into the flat buffers (and the per-parameter tensors released)
    assert all(p_.grad is None for _, p_ in tr.g_flat.params)
    assert tr.g_flat.grad.abs().sum().item() > 0 and tr.d_flat.grad.abs().sum().item() > 0
    tr.ema_update()
    assert torch.isfinite(tr.ema_flat.data).all()
    # lazy regularisers run
    tr.d_regularize(real.to(DEV))
    tr._latents = lambda n: (torch.randn(n, 512, 16, device=DEV), torch.randn(n, 512, 16, device=DEV))
    tr.g_regularize()
    assert torch.isfinite(tr.g_flat.data).all() and torch.isfinite(tr.d_flat.data).all()
    out = tr.step_from_host((torch.rand(4, 3, 32, 32) * 2 - 1).pin_memory())
    assert all(np.isfinite(v) for v in out.values())


def test_generator_1024_matches_oracle_fp32():
    """BASELINE configs[3] shape (1024^2 generator, 32/64-channel tail layers) in fp32 parity mode against
    the CPU oracle with the same deterministic weights (batch 1 keeps the CPU side to a few seconds)."""
    import model_spatial_query as M
    size, cm = 1024, 2
    sdg = O.synthetic_state(O.generator_shapes(size, cm))
    g = M.Generator(size, 512, 512, 18, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1)
    g.load_state_dict(sdg, strict=True)
    g = g.to(DEV)
    gen = torch.Generator().manual_seed(5)
    z, p = torch.randn(1, 512, 16, generator=gen), torch.randn(1, 512, 16, generator=gen)
    with torch.no_grad():
        img, lat, _ = g(z.to(DEV), p.to(DEV), return_latents=True)
        ref, lat_ref = O.generator_forward(sdg, z, p, size)
    assert img.shape == (1, 3, 1024, 1024) and lat.shape == (1, 18, 512)
    assert (lat.cpu() - lat_ref).abs().max().item() < 1e-3
    assert (img.cpu() - ref).abs().max().item() < 1e-3


def test_checkpoint_resume_reproduces_the_next_iteration(tmp_path):
    """save_checkpoint / load_checkpoint in the reference's {'g','d','g_ema','g_optim','d_optim'} layout: a
    second trainer resumed from the file takes the same next iteration; the optimiser entries load into
    torch.optim.Adam (what the reference's resume code does, train_spatial_query.py:487-492)."""
    from transeditor_b200.checkpoint import load_checkpoint, save_checkpoint
    from transeditor_b200.train_step import TrainConfig, Trainer
    cfg = TrainConfig(size=32, batch=4)
    a = Trainer(cfg, DEV, seed=0)
    real = (torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(7)) * 2 - 1).to(DEV)
    for _ in range(2):
        a.step(real)
    path = tmp_path / "000002.pt"
    save_checkpoint(a, path)
    ckpt = torch.load(path, map_location="cpu")
    assert sorted(ckpt) == ["d", "d_optim", "g", "g_ema", "g_optim"]
    torch.optim.Adam(a.discriminator.parameters(), lr=1.0).load_state_dict(ckpt["d_optim"])
    torch.optim.Adam(a.generator.parameters(), lr=1.0).load_state_dict(ckpt["g_optim"])
    b = Trainer(cfg, DEV, seed=5)
    load_checkpoint(b, path)
    b.iteration = a.iteration
    b.mean_path_length.copy_(a.mean_path_length)
    for t in (a, b):
        torch.manual_seed(4242)
        t.step(real)
    # split-K atomics make the gradients order-dependent in the last bits and Adam's lr*g/(|g|+eps) is
    # ill-conditioned where g ~ eps: compare like test_train_step_matches_cpu_oracle_update does
    for x, y in ((a.g_flat.data, b.g_flat.data), (a.d_flat.data, b.d_flat.data), (a.ema_flat.data, b.ema_flat.data)):
        diff = (x - y).abs()
        assert diff.mean().item() < 1e-5 and (diff > 2e-4).float().mean().item() < 1e-3
    # an un-resumed trainer is far away: the comparison above is not vacuous
    c = Trainer(cfg, DEV, seed=5)
    assert (a.g_flat.data - c.g_flat.data).abs().mean().item() > 1e-2
