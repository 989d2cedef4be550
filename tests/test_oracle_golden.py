"""Pins oracle/te_oracle.py and oracle/ops_cpu.py against golden vectors produced by the
UNMODIFIED reference classes (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ops_cpu, te_oracle as O
from tests.conftest import load_golden, small


def _state(size, cm):
    return (O.synthetic_state(O.generator_shapes(size, cm)),
            O.synthetic_state(O.discriminator_shapes(size, cm)))


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


@pytest.mark.parametrize("name", ["gd32_b4", "gd64_b2", "gd256_b1"])
def test_forward_matches_reference(name):
    gold = load_golden(name)
    size, cm = int(gold["size"]), int(gold["cm"])
    sdg, sdd = _state(size, cm)
    # the weights are regenerated, not stored: make sure they are the ones the golden run used
    assert abs(_checksum(sdg) - float(gold["g_checksum"])) < 1e-6 * float(gold["g_checksum"])
    assert abs(_checksum(sdd) - float(gold["d_checksum"])) < 1e-6 * float(gold["d_checksum"])
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    with torch.no_grad():
        img, lat = O.generator_forward(sdg, z, p, size)
        assert np.abs(img.numpy() - gold["img"]).max() < 1e-4
        assert np.abs(lat.numpy() - gold["latent"]).max() < 1e-4
        lat2, pp = O.generator_front(sdg, z, p)
        assert np.abs(pp.numpy() - gold["p_plus"]).max() < 1e-5
        img2, _ = O.generator_forward(sdg, torch.from_numpy(gold["z_plus"]), torch.from_numpy(gold["p_plus"]),
                                      size, use_spatial_mapping=False, use_style_mapping=False)
        assert np.abs(img2.numpy() - gold["img_from_plus"]).max() < 1e-4
        assert np.abs(O.discriminator_forward(sdd, img).numpy() - gold["d_fake"]).max() < 1e-4
        real = torch.from_numpy(gold["real"])
        assert np.abs(O.discriminator_forward(sdd, real).numpy() - gold["d_real"]).max() < 1e-4


def test_gradients_and_regularisers_match_reference():
    gold = load_golden("gd32_b4")
    sdg, sdd = _state(32, 2)
    gk = [k[6:] for k in gold if k.startswith("ggrad.")]
    dk = [k[13:] for k in gold if k.startswith("dgrad_from_g.")]
    for k in gk:
        sdg[k].requires_grad_(True)
    for k in dk:
        sdd[k].requires_grad_(True)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    img, lat = O.generator_forward(sdg, z, p, 32)
    loss = O.g_nonsaturating_loss(O.discriminator_forward(sdd, img))
    loss.backward()
    assert abs(float(loss.detach()) - float(gold["g_loss"])) < 1e-5
    for k in gk:
        v = gold["ggrad." + k]
        assert np.abs(small(sdg[k].grad) - v).max() < 1e-4 * max(1.0, np.abs(v).max()), k
    for k in dk:
        v = gold["dgrad_from_g." + k]
        assert np.abs(small(sdd[k].grad) - v).max() < 1e-4 * max(1.0, np.abs(v).max()), k
    # R1
    sdg, sdd = _state(32, 2)
    rk = [k[7:] for k in gold if k.startswith("r1grad.")]
    for k in rk:
        sdd[k].requires_grad_(True)
    real = torch.from_numpy(gold["real"]).requires_grad_(True)
    pen = O.d_r1_penalty(O.discriminator_forward(sdd, real), real)
    pen.backward()
    assert abs(float(pen.detach()) - float(gold["r1"])) < 1e-4 * max(1.0, float(gold["r1"]))
    for k in rk:
        v = gold["r1grad." + k]
        assert np.abs(small(sdd[k].grad) - v).max() < 1e-4 * max(1.0, np.abs(v).max()), k
    # path length
    pk = [k[9:] for k in gold if k.startswith("pathgrad.")]
    for k in pk:
        sdg[k].requires_grad_(True)
    img, lat = O.generator_forward(sdg, z, p, 32)
    pl = O.g_path_lengths(img, lat, torch.from_numpy(gold["path_noise"]))
    (pl - 0.5).pow(2).mean().backward()
    assert np.abs(pl.detach().numpy() - gold["path_lengths"]).max() < 1e-4 * max(1.0, np.abs(gold["path_lengths"]).max())
    for k in pk:
        v = gold["pathgrad." + k]
        assert np.abs(small(sdg[k].grad) - v).max() < 1e-4 * max(1.0, np.abs(v).max()), k


def test_noise_route_matches_reference():
    gold = load_golden("g32_noise")
    sdg, _ = _state(32, 2)
    z, p = torch.from_numpy(gold["z"]), torch.from_numpy(gold["p"])
    noise = [torch.from_numpy(gold[f"noise_{i}"]) for i in range(7)]
    with torch.no_grad():
        img, _ = O.generator_forward(sdg, z, p, 32, noise=noise, inject_noise=True)
        assert np.abs(img.numpy() - gold["img"]).max() < 1e-4
        buf = [sdg[f"noises.noise_{i}"] for i in range(7)]
        img, _ = O.generator_forward(sdg, z, p, 32, noise=buf, inject_noise=True)
        assert np.abs(img.numpy() - gold["img_buffer_noise"]).max() < 1e-4


def test_ops_match_kernel_index_transcription():
    """ops.npz holds outputs of the literal transcription of the reference CUDA kernel's index
    arithmetic (asymmetric FIR, every live up/down/pad set); the conv-based restatement must agree."""
    gold = load_golden("ops")
    fir = torch.from_numpy(gold["fir"])
    for ci, (up, down, p0, p1) in enumerate(gold["cases"]):
        for hw in ("5x7", "8x8", "9x6"):
            x = torch.from_numpy(gold[f"up_x_{ci}_{hw}"])
            y = ops_cpu.upfirdn2d_planes(x[None], fir, int(up), int(up), int(down), int(down),
                                         int(p0), int(p1), int(p0), int(p1))[0]
            ref = gold[f"up_y_{ci}_{hw}"]
            assert y.shape == ref.shape
            assert np.abs(y.numpy() - ref).max() < 1e-12
    x, b = torch.from_numpy(gold["fba_x"]), torch.from_numpy(gold["fba_b"])
    assert np.abs(ops_cpu.fused_leaky_relu(x, b).numpy() - gold["fba_30"]).max() < 1e-6


def test_product_attention_restatement_matches_the_oracle():
    """`op.attn_stack_reference` (the differentiable re-expression the product uses for second-order gradients of the
    fused attention stack) computes the oracle's `interaction_stack` — which the golden tests above pin to the
    reference's own AttentionBlock."""
    import torch
    from oracle import te_oracle as O
    from transeditor_b200 import op
    g = torch.Generator().manual_seed(21)
    lr = 0.01
    blocks = []
    for i in range(3):
        d = 528 if i == 0 else 512
        w = lambda o, n: (torch.randn(o, n, generator=g, dtype=torch.float64) / lr)  # noqa: E731
        b = lambda o: (torch.randn(o, generator=g, dtype=torch.float64) * 0.3 / lr)  # noqa: E731
        blocks.append({"in_dim": d, "param_dim": d, "w_proj": w(512, d) if i == 0 else None,
                       "b_proj": b(512) if i == 0 else None, "w_q": w(128, d), "b_q": b(128), "w_k": w(128, d),
                       "b_k": b(128), "w_v": w(128, d), "b_v": b(128), "w_o": w(512, 128), "b_o": b(512),
                       "w_m1": w(512, 512), "b_m1": b(512), "w_m2": w(512, 512), "b_m2": b(512)})
    x0 = torch.randn(2, 16, 528, generator=g, dtype=torch.float64)
    p0 = torch.randn(2, 16, 528, generator=g, dtype=torch.float64)
    p = torch.randn(2, 16, 512, generator=g, dtype=torch.float64)
    a = op.attn_stack_reference(x0, p0, p, blocks, lr)
    r = O.interaction_stack(x0, p0, p, blocks)
    assert (a - r).abs().max().item() < 1e-10 * max(1.0, r.abs().max().item())
