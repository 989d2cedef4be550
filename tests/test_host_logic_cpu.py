"""Host-side logic on CPU: the product's Python layer (autograd wiring, conv geometry algebra,
Generator/Discriminator orchestration and flag routing) driven through tests/emu.py and checked
against the golden vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

from oracle import te_oracle as O
from tests.conftest import load_golden, small

pytestmark = pytest.mark.usefixtures("cpu_emulation")


def _models(size, cm, inject_noise=False):
    import model_spatial_query as M
    t = 2 * int(np.log2(size)) - 2
    g = M.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1,
                    layer_noise_injection=inject_noise)
    d = M.Discriminator(size, channel_multiplier=cm)
    g.load_state_dict(O.synthetic_state(O.generator_shapes(size, cm)), strict=True)
    d.load_state_dict(O.synthetic_state(O.discriminator_shapes(size, cm)), strict=True)
    return g, d


def test_forward_and_routes_match_reference_golden():
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    with torch.no_grad():
        img, lat, none = g(z, p, return_latents=True)
        assert none is None
        assert np.abs(img.numpy() - gold["img"]).max() < 1e-3
        assert np.abs(lat.numpy() - gold["latent"]).max() < 1e-3
        zp, pp = g(z, p, return_mapped_codes=True)
        assert np.abs(zp.numpy() - gold["z_plus"]).max() < 1e-4
        assert np.abs(pp.numpy() - gold["p_plus"]).max() < 1e-4
        assert torch.equal(g(z, p, return_only_mapped_p=True), pp)
        assert torch.equal(g(z, p, return_only_mapped_z=True), zp)
        img2, a, b = g(zp, pp, use_spatial_mapping=False, use_style_mapping=False)
        assert a is None and b is None
        assert np.abs(img2.numpy() - gold["img_from_plus"]).max() < 1e-3
        assert np.abs(d(img).numpy() - gold["d_fake"]).max() < 1e-3
        assert np.abs(d(torch.from_numpy(gold["real"])).numpy() - gold["d_real"]).max() < 1e-3
        # one stacked pass over [fake; real] = the two separate calls (per-sub-batch minibatch stddev)
        both = d.forward_stacked(torch.cat([img, torch.from_numpy(gold["real"])]), 2)
        assert np.abs(both[:4].numpy() - gold["d_fake"]).max() < 1e-3
        assert np.abs(both[4:].numpy() - gold["d_real"]).max() < 1e-3
        only_lat = g(z, p, return_only_style_latent=True)
        assert torch.allclose(only_lat, lat)
        im3, l3 = g(z, p, return_style=True)
        assert torch.allclose(im3, img) and torch.allclose(l3, lat)
        im4, pl4 = g(z, p, return_p_latent=True)
        assert pl4.shape == (4, 16, 512)


def test_generator_step_gradients_match_reference():
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    img, lat, _ = g(z, p, return_latents=True)
    loss = torch.nn.functional.softplus(-d(img)).mean()
    loss.backward()
    assert abs(float(loss) - float(gold["g_loss"])) < 1e-4
    gp, dp = dict(g.named_parameters()), dict(d.named_parameters())
    for key, val in gold.items():
        if key.startswith("ggrad."):
            got = small(gp[key[6:]].grad)
        elif key.startswith("dgrad_from_g."):
            got = small(dp[key[13:]].grad)
        else:
            continue
        tol = 2e-3 * max(1.0, float(np.abs(val).max()))
        assert np.abs(got - val).max() < tol, key


def test_r1_and_path_regularisers_match_reference():
    gold = load_golden("gd32_b4")
    g, d = _models(32, 2)
    dp, gp = dict(d.named_parameters()), dict(g.named_parameters())
    real = torch.from_numpy(gold["real"]).requires_grad_(True)
    pred = d(real)
    (gi,) = torch.autograd.grad(pred.sum(), real, create_graph=True)
    r1 = gi.pow(2).reshape(gi.shape[0], -1).sum(1).mean()
    r1.backward()
    assert abs(float(r1) - float(gold["r1"])) < 1e-3 * max(1.0, abs(float(gold["r1"])))
    assert np.abs(gi.detach().numpy() - gold["r1_grad_img"]).max() < 1e-3 * max(1.0, np.abs(gold["r1_grad_img"]).max())
    for key, val in gold.items():
        if key.startswith("r1grad."):
            got = small(dp[key[7:]].grad)
            assert np.abs(got - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    img, lat, _ = g(z, p, return_latents=True)
    noise = torch.from_numpy(gold["path_noise"])
    (gl,) = torch.autograd.grad((img * noise).sum(), lat, create_graph=True)
    pl = torch.sqrt(gl.pow(2).sum(2).mean(1))
    (pl - 0.5).pow(2).mean().backward()
    assert np.abs(pl.detach().numpy() - gold["path_lengths"]).max() < 1e-3 * max(1.0, np.abs(gold["path_lengths"]).max())
    for key, val in gold.items():
        if key.startswith("pathgrad."):
            got = small(gp[key[9:]].grad)
            assert np.abs(got - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key


def test_noise_injection_routes():
    gold = load_golden("g32_noise")
    g, _ = _models(32, 2, inject_noise=True)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    noise = [torch.from_numpy(gold[f"noise_{i}"]) for i in range(7)]
    with torch.no_grad():
        img, _, _ = g(z, p, noise=noise)
        assert np.abs(img.numpy() - gold["img"]).max() < 1e-3
        img_b, _, _ = g(z, p, randomize_noise=False)
        assert np.abs(img_b.numpy() - gold["img_buffer_noise"]).max() < 1e-3
    # differentiable (unfused) route gives the same image as the fused inference route
    zz = z.clone().requires_grad_(True)
    img_g, _, _ = g(zz, p, noise=noise)
    assert np.abs(img_g.detach().numpy() - gold["img"]).max() < 1e-3


def test_pack_cache_host_logic(cpu_emulation):
    """tc.PackCache (weights repacked once per optimiser step): hits only for registered whole parameters, repacks on
    torch-side version bumps, refresh() after raw-pointer updates; emulated table-pack kernel == per-call repack."""
    from transeditor_b200 import tc
    g = torch.Generator().manual_seed(0)
    w = torch.nn.Parameter(torch.randn(24, 40, 3, 3, generator=g))
    w1 = torch.nn.Parameter(torch.randn(1, 8, 16, 1, 1, generator=g))   # ModulatedConv2d keeps a leading 1
    cache = tc.PackCache()
    cache.register([w, w1])
    tc.set_pack_cache(None)
    plain = tc.pack_weight(w.detach(), False, 0.25), tc.pack_weight(w.detach(), True, 0.25)
    tc.set_pack_cache(cache)
    torch.set_grad_enabled(False)   # pack_weight runs inside autograd.Function.forward in the product
    try:
        a, at = tc.pack_weight(w, False, 0.25), tc.pack_weight(w, True, 0.25)
        assert torch.equal(a, plain[0]) and torch.equal(at, plain[1])
        assert tc.pack_weight(w, False, 0.25) is a and tc.pack_weight(w[:, :8], False, 0.25) is not a
        assert tc.pack_weight(w1[0], False, 1.0) is tc.pack_weight(w1[0], False, 1.0)
        w.mul_(3.0)                                       # version bump -> both orientations repacked on next use
        assert torch.equal(tc.pack_weight(w, True, 0.25).float(), (plain[1].float() * 3).to(torch.bfloat16).float()) or \
            torch.allclose(tc.pack_weight(w, True, 0.25).float(), plain[1].float() * 3, rtol=1e-2)
        w.data.add_(1.0)                                  # raw update, no version bump: stale until refresh()
        lo = w.data_ptr()
        cache.refresh(lo, lo + 1)
        tc.set_pack_cache(None)
        fresh = tc.pack_weight(w.detach(), False, 0.25)
        assert torch.equal(a, fresh)
    finally:
        torch.set_grad_enabled(True)
        tc.set_pack_cache(None)


@pytest.mark.parametrize("space", ["p", "p_plus"])
def test_spatial_path_regulariser_matches_reference(space):
    """train_spatial_query.py:252-277 in both spaces: second order through the spatial mapping network, the
    cross-attention stack (fused route: AttnStack re-expresses its backward differentiably) and the const-input
    convolutions, against the reference's own numbers."""
    gold = load_golden("gd32_spatial_path")
    g, _ = _models(32, 2)
    gp = dict(g.named_parameters())
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    noise = torch.from_numpy(gold["noise"])
    if space == "p":
        target = p.clone().requires_grad_()
        img, _, _ = g(z, target)
    else:
        target = g(z, p, return_only_mapped_p=True)
        target.requires_grad_()
        img, _, _ = g(z, target, use_spatial_mapping=False)
    (gl,) = torch.autograd.grad((img * noise).sum(), target, create_graph=True)
    pl = torch.sqrt(gl.pow(2).sum(2).mean(1))
    (pl - 0.25).pow(2).mean().backward()
    ref = gold[space + ".lengths"]
    assert np.abs(pl.detach().numpy() - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
    seen = 0
    for key, val in gold.items():
        if key.startswith(space + ".grad."):
            got = small(gp[key[len(space) + 6:]].grad)
            assert np.abs(got - val).max() < 2e-3 * max(1.0, float(np.abs(val).max())), key
            seen += 1
    assert seen >= 4


def test_path_regulariser_steps_rgb_biases_like_torch_adam(cpu_emulation):
    """the reference adds `0 * fake_img[0,0,0,0]` to the path penalty (train_spatial_query.py:240-243), so
    the to_rgb biases get a ZERO gradient (not None) and torch.optim.Adam steps them: step += 1, exp_avg_sq *= beta2,
    parameter unchanged (beta1 = 0).  g_step -> g_regularize -> g_step must reproduce that bookkeeping; the noise
    strengths (grad None in the reference) stay untouched."""
    from transeditor_b200.train_step import TrainConfig, Trainer
    tr = Trainer(TrainConfig(size=32, batch=4), "cpu", seed=0)
    flat, opt = tr.g_flat, tr.g_optim
    lo, hi = flat.group_end[0], flat.group_end[1]          # group 1 = the to_rgb biases
    assert hi - lo >= 4 * 3 and flat.group_end[2] > hi     # 4-aligned slots of 3 values; group 2 = noise strengths
    tr.g_step()
    assert int(opt.steps[0]) == 1 and int(opt.steps[1]) == 1 and int(opt.steps[2]) == 0
    p1, v1 = flat.data[lo:hi].clone(), opt.v[lo:hi].clone()
    assert v1.abs().sum() > 0
    tr.g_regularize()
    assert int(opt.steps[0]) == 2 and int(opt.steps[1]) == 2 and int(opt.steps[2]) == 0
    assert torch.equal(flat.data[lo:hi], p1)                              # zero gradient, beta1 = 0: no move
    assert torch.allclose(opt.v[lo:hi], v1 * opt.betas[1], rtol=1e-6)     # second moment decays
    # reference run of the same three updates on one bias with torch.optim.Adam
    name = next(n for n, _ in flat.params if n == "to_rgb1.bias")
    o = flat.offsets[name]
    g1 = None
    ref_p = torch.nn.Parameter(torch.zeros(3))
    ref_opt = torch.optim.Adam([ref_p], lr=opt.lr, betas=opt.betas, eps=opt.eps)
    v_after_1 = opt.v[o:o + 3].clone() / opt.betas[1]   # undo the decay of the regulariser step
    g1 = (v_after_1 / (1 - opt.betas[1])).sqrt()          # |g| of the first step (first moment of g^2)
    ref_p.grad = g1.clone()
    ref_opt.step()
    ref_p.grad = torch.zeros(3)
    ref_opt.step()
    st = ref_opt.state[ref_p]
    assert int(st["step"]) == 2
    assert torch.allclose(st["exp_avg_sq"], opt.v[o:o + 3], rtol=1e-5, atol=1e-20)
    tr.g_step()
    assert int(opt.steps[1]) == 3 and int(opt.steps[2]) == 0


def test_ema_rides_in_the_last_generator_update(cpu_emulation):
    """step() fuses accumulate(g_ema, g, decay) (train_spatial_query.py:56-61,294) into the optimiser launch of the
    iteration's LAST generator phase: same numbers as the stand-alone pass."""
    from transeditor_b200.train_step import TrainConfig, Trainer
    real = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(1)) * 2 - 1
    a = Trainer(TrainConfig(size=32, batch=4), "cpu", seed=0)
    b = Trainer(TrainConfig(size=32, batch=4), "cpu", seed=0)
    for it in range(2):   # iteration 0 ends with the path regulariser, iteration 1 with the plain G step
        torch.manual_seed(77 + it)
        a.step(real)
        torch.manual_seed(77 + it)
        b.d_step(real)
        if it % 16 == 0:
            b.d_regularize(real)
        b.g_step()
        if it % 4 == 0:
            b.g_regularize()
        b.ema_update()
        b.iteration += 1
    assert torch.equal(a.g_flat.data, b.g_flat.data)
    assert torch.allclose(a.ema_flat.data, b.ema_flat.data, rtol=1e-6, atol=1e-7)
    assert not torch.equal(a.ema_flat.data, a.g_flat.data)
