"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes on CPU.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The vectors are small; the weights are NOT stored — they are regenerated
deterministically from key names by oracle.te_oracle.synthetic_state (same torch
build on the GPU box), and the golden file records a checksum of them.

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, te_oracle as O  # noqa: E402
from oracle import ops_cpu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _inputs(seed, b):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 512, 16, generator=g), torch.randn(b, 512, 16, generator=g)


def _small(t):
    """Large gradients are stored as a strided sample (stride 97, <= 8192 values);
    tests apply the same sampling to the product's gradient."""
    a = t.detach().numpy().reshape(-1)
    return a[::97][:8192].copy() if a.size > 65536 else a.copy().reshape(t.shape)


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_models(ref, size, cm, inject_noise=False, n_trans=8):
    t = 2 * int(np.log2(size)) - 2
    with ref_shim.cpu_mode():
        g = ref.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=n_trans,
                          pixel_norm_op_dim=1, layer_noise_injection=inject_noise).eval()
        d = ref.Discriminator(size, channel_multiplier=cm).eval()
    sdg = O.synthetic_state(O.generator_shapes(size, cm, n_trans))
    sdd = O.synthetic_state(O.discriminator_shapes(size, cm))
    g.load_state_dict(sdg, strict=True)
    d.load_state_dict(sdd, strict=True)
    return g, d, sdg, sdd


def golden_model(ref, name, size, cm, b, seed, with_grads):
    g, d, sdg, sdd = build_models(ref, size, cm)
    z, p = _inputs(seed, b)
    rec = {"size": size, "cm": cm, "z": z.numpy(), "p": p.numpy(),
           "g_checksum": _checksum(sdg), "d_checksum": _checksum(sdd)}
    with ref_shim.cpu_mode():
        with torch.no_grad():
            img, lat, _ = g(z, p, return_latents=True)
            rec["img"] = img.numpy()
            rec["latent"] = lat.numpy()
            zp, pp = g(z, p, return_mapped_codes=True)
            rec["z_plus"] = zp.numpy()
            rec["p_plus"] = pp.numpy()
            img2, _, _ = g(zp, pp, use_spatial_mapping=False, use_style_mapping=False)
            rec["img_from_plus"] = img2.numpy()
            rec["d_fake"] = d(img).numpy()
            gr = torch.Generator().manual_seed(seed + 1)
            real = torch.rand(b, 3, size, size, generator=gr) * 2 - 1
            rec["real"] = real.numpy()
            rec["d_real"] = d(real).numpy()
        if with_grads:
            # G step: non-saturating loss through D into G (train_spatial_query.py:210-224)
            for q in list(g.parameters()) + list(d.parameters()):
                q.grad = None
            img, lat, _ = g(z, p, return_latents=True)
            loss = torch.nn.functional.softplus(-d(img)).mean()
            loss.backward()
            rec["g_loss"] = loss.detach().numpy()
            gp = dict(g.named_parameters())
            for k in ("to_rgb1.bias", "conv1.activate.bias", "adjust_style.weight",
                      "convs.0.conv.modulation.bias", "interact.0.atten.q_transform.bias",
                      "convs.1.conv.weight", "convs.0.conv.weight", "to_rgbs.0.conv.weight"):
                rec["ggrad." + k] = _small(gp[k].grad)
            dp = dict(d.named_parameters())
            for k in ("final_linear.1.weight", "convs.0.1.bias", "convs.1.conv2.1.weight",
                      "convs.1.skip.1.weight"):
                rec["dgrad_from_g." + k] = _small(dp[k].grad)
            # R1 (train_spatial_query.py:196-206)
            for q in d.parameters():
                q.grad = None
            real_r = real.clone().requires_grad_(True)
            pred = d(real_r)
            (gi,) = torch.autograd.grad(pred.sum(), real_r, create_graph=True)
            r1 = gi.pow(2).reshape(b, -1).sum(1).mean()
            r1.backward()
            rec["r1"] = r1.detach().numpy()
            rec["r1_grad_img"] = gi.detach().numpy()
            for k in ("final_linear.1.weight", "convs.0.1.bias", "convs.1.conv2.1.weight",
                      "convs.0.0.weight"):
                rec["r1grad." + k] = _small(dp[k].grad)
            # path length (train_spatial_query.py:92-105), explicit projection noise
            for q in g.parameters():
                q.grad = None
            img, lat, _ = g(z, p, return_latents=True)
            gn = torch.Generator().manual_seed(seed + 2)
            noise = torch.randn(img.shape, generator=gn) / np.sqrt(size * size)
            rec["path_noise"] = noise.numpy()
            (gl,) = torch.autograd.grad((img * noise).sum(), lat, create_graph=True)
            pl = torch.sqrt(gl.pow(2).sum(2).mean(1))
            pen = (pl - 0.5).pow(2).mean()
            pen.backward()
            rec["path_lengths"] = pl.detach().numpy()
            for k in ("conv1.activate.bias", "adjust_style.weight", "convs.0.conv.modulation.bias",
                      "convs.1.conv.weight", "conv1.conv.modulation.weight"):
                rec["pathgrad." + k] = _small(gp[k].grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "img std", float(rec["img"].std()), "bytes",
          os.path.getsize(os.path.join(OUT, name + ".npz")))


def golden_noise(ref):
    """inject_noise=True route with explicit per-layer noise (model_spatial_query.py:397-399)."""
    size, cm, b = 32, 2, 2
    g, _, sdg, _ = build_models(ref, size, cm, inject_noise=True)
    z, p = _inputs(77, b)
    gn = torch.Generator().manual_seed(78)
    noise = [torch.randn(1, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), generator=gn) for i in range(7)]
    with ref_shim.cpu_mode(), torch.no_grad():
        img, _, _ = g(z, p, noise=noise)
        img_buf, _, _ = g(z, p, randomize_noise=False)
    rec = {"size": size, "cm": cm, "z": z.numpy(), "p": p.numpy(), "img": img.numpy(),
           "img_buffer_noise": img_buf.numpy(), "g_checksum": _checksum(sdg)}
    for i, n in enumerate(noise):
        rec[f"noise_{i}"] = n.numpy()
    np.savez_compressed(os.path.join(OUT, "g32_noise.npz"), **rec)
    print("g32_noise", float(img.std()))


def golden_spatial_path(ref):
    """The optional spatial path regulariser (train_spatial_query.py:252-277) in both of its spaces, with an explicit
    projection noise: path lengths and a few parameter gradients of the penalty (second order through the mapping
    network and the cross-attention stack)."""
    size, cm, b = 32, 2, 2
    g, _, sdg, _ = build_models(ref, size, cm)
    z, p = _inputs(91, b)
    gn = torch.Generator().manual_seed(92)
    noise = torch.randn(b, 3, size, size, generator=gn) / np.sqrt(size * size)
    rec = {"size": size, "cm": cm, "z": z.numpy(), "p": p.numpy(), "noise": noise.numpy(), "g_checksum": _checksum(sdg)}
    gp = dict(g.named_parameters())
    keys = ("spatial_mapping_network.1.weight", "interact.0.atten.q_transform.weight", "interact.3.mlp.0.bias",
            "conv1.conv.weight", "adjust_style.weight", "convs.0.conv.modulation.bias")
    with ref_shim.cpu_mode():
        for space in ("p", "p_plus"):
            for q in g.parameters():
                q.grad = None
            if space == "p":                       # :258-265
                target = p.clone().requires_grad_()
                img, _, _ = g(z, target)
            else:                                  # :266-272
                target = g(z, p, return_only_mapped_p=True)
                target.requires_grad_()
                img, _, _ = g(z, target, use_spatial_mapping=False)
            (gl,) = torch.autograd.grad((img * noise).sum(), target, create_graph=True)
            pl = torch.sqrt(gl.pow(2).sum(2).mean(1))          # g_path_regularize, :92-105
            (pl - 0.25).pow(2).mean().backward()
            rec[space + ".lengths"] = pl.detach().numpy()
            for k in keys:
                if gp[k].grad is not None:
                    rec[space + ".grad." + k] = _small(gp[k].grad)
    np.savez_compressed(os.path.join(OUT, "gd32_spatial_path.npz"), **rec)
    print("gd32_spatial_path", {k: v.shape for k, v in rec.items() if k.endswith("lengths")})


def golden_ops():
    """Operator-level vectors.  upfirdn2d: every live parameter set with an
    ASYMMETRIC random FIR, produced by the literal kernel-index transcription
    (ops_cpu.upfirdn2d_index_form, upfirdn2d_kernel.cu:84-133); fused_bias_act: all defined
    (act, grad) cases of fused_bias_act_kernel.cu:36-45."""
    g = torch.Generator().manual_seed(5)
    rec = {}
    fir = torch.rand(4, 4, generator=g, dtype=torch.float64)
    rec["fir"] = fir.numpy()
    cases = [(1, 1, 1, 1), (1, 1, 2, 2), (2, 1, 2, 1), (1, 2, 1, 1), (1, 2, 2, 2), (1, 1, 0, 0),
             (2, 2, 1, 1), (1, 1, 3, 0), (2, 1, 1, 2)]
    rec["cases"] = np.array(cases)
    for ci, (up, down, p0, p1) in enumerate(cases):
        for h, w in ((5, 7), (8, 8), (9, 6)):
            x = torch.randn(h, w, generator=g, dtype=torch.float64)
            y = ops_cpu.upfirdn2d_index_form(x, fir, up, down, p0, p1)
            rec[f"up_x_{ci}_{h}x{w}"] = x.numpy()
            rec[f"up_y_{ci}_{h}x{w}"] = y.numpy()
    x = torch.randn(3, 5, 4, 6, generator=g)
    bias = torch.randn(5, generator=g)
    ref = torch.randn(3, 5, 4, 6, generator=g)
    rec["fba_x"], rec["fba_b"], rec["fba_ref"] = x.numpy(), bias.numpy(), ref.numpy()
    xb = x + bias.view(1, -1, 1, 1)
    a, s = 0.2, 2 ** 0.5
    rec["fba_30"] = (torch.where(xb > 0, xb, xb * a) * s).numpy()
    rec["fba_31"] = (torch.where(ref > 0, xb, xb * a) * s).numpy()
    rec["fba_31_nobias"] = (torch.where(ref > 0, x, x * a) * s).numpy()
    rec["fba_10"] = (xb * s).numpy()
    rec["fba_32"] = (xb * 0).numpy()
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **rec)
    print("ops done")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_shim.load_reference_module()
    golden_ops()
    golden_model(ref, "gd32_b4", 32, 2, 4, 11, with_grads=True)
    golden_model(ref, "gd64_b2", 64, 1, 2, 21, with_grads=False)
    golden_model(ref, "gd256_b1", 256, 2, 1, 31, with_grads=False)
    golden_noise(ref)
    golden_spatial_path(ref)


if __name__ == "__main__":
    main()
