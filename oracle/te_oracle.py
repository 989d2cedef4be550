"""Functional CPU restatement of the TransEditor generator / discriminator hot path.

TEST INFRASTRUCTURE ONLY — the checker, never the thing measured or shipped.  Only
tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import it; nothing under transeditor_b200/ does.

Parity status: PINNED.  The reference has no golden vectors of its own,
so this file is pinned against outputs of the reference's own classes run in the build
container (oracle/ref_shim.py -> oracle/make_golden.py -> tests/golden/*.npz), see
tests/test_oracle_golden.py.

Design: no nn.Module.  Every function takes a plain `state_dict`-style mapping with
the reference's key names plus the tensors, and is written with
differentiable torch ops in the reference's *literal* formulation (per-sample weight
materialisation + grouped convolution), which is deliberately different from the
product's shared-weight formulation so the two cannot share a bug.

Line citations are into /root/reference/model_spatial_query.py unless noted.
"""
import math

import torch
import torch.nn.functional as F

from oracle.ops_cpu import fused_leaky_relu, upfirdn2d

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------- helpers
def fir_kernel(taps=(1, 3, 3, 1), gain=1.0):
    """:84-92 make_kernel — outer product of the 1-D taps, normalised to sum 1."""
    k = torch.tensor(taps, dtype=torch.float32)
    k = k[None, :] * k[:, None]
    return k / k.sum() * gain


def pixel_norm(x, dim):
    """:75-81"""
    return x * torch.rsqrt(torch.mean(x * x, dim=dim, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """:194-221.  scale = lr_mul / sqrt(in_dim); bias is multiplied by lr_mul;
    with activation the bias goes through fused_leaky_relu."""
    scale = lr_mul / math.sqrt(weight.shape[1])
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, None if bias is None else bias * lr_mul)


def sub(sd, prefix):
    """View of the entries below `prefix.` with the prefix stripped."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


# --------------------------------------------------------------------------- generator
def modulated_conv(x, style_vec, p, demodulate, upsample, fir_up):
    """:296-337.  `p` holds weight[1,O,I,k,k], modulation.weight, modulation.bias."""
    w = p["weight"]
    _, cout, cin, k, _ = w.shape
    b, _, h, wd = x.shape
    s = equal_linear(style_vec, p["modulation.weight"], p["modulation.bias"])  # :299
    wb = (1.0 / math.sqrt(cin * k * k)) * w * s.view(b, 1, cin, 1, 1)  # :300
    if demodulate:
        d = torch.rsqrt(wb.pow(2).sum([2, 3, 4]) + 1e-8)  # :303
        wb = wb * d.view(b, cout, 1, 1, 1)
    if upsample:
        xg = x.reshape(1, b * cin, h, wd)
        wt = wb.transpose(1, 2).reshape(b * cin, cout, k, k)  # :315-317
        y = F.conv_transpose2d(xg, wt, padding=0, stride=2, groups=b)  # :318
        y = y.view(b, cout, y.shape[2], y.shape[3])
        # Blur(pad=(pad0,pad1), FIR*factor^2) with factor 2, :262-268,321
        pp = (fir_up.shape[0] - 2) - (k - 1)
        pad = ((pp + 1) // 2 + 1, pp // 2 + 1)
        return upfirdn2d(y, fir_up, pad=pad)
    xg = x.reshape(1, b * cin, h, wd)
    y = F.conv2d(xg, wb.view(b * cout, cin, k, k), padding=k // 2, groups=b)  # :333
    return y.view(b, cout, y.shape[2], y.shape[3])


def styled_conv(x, style_vec, p, upsample, fir_up, noise, inject_noise):
    """:395-403.  NoiseInjection (:346-351) only runs when layer_noise_injection."""
    y = modulated_conv(x, style_vec, sub(p, "conv"), True, upsample, fir_up)
    if inject_noise:
        if noise is None:
            noise = torch.randn(y.shape[0], 1, y.shape[2], y.shape[3], dtype=y.dtype)
        y = y + p["noise.weight"] * noise
    return fused_leaky_relu(y, p["activate.bias"])


def to_rgb(x, style_vec, p, skip, fir_up):
    """:416-425.  1x1 modconv without demodulation + bias + upsampled skip."""
    y = modulated_conv(x, style_vec, sub(p, "conv"), False, False, None) + p["bias"]
    if skip is not None:
        # Upsample: FIR*4, pad (2,1) for the 4-tap FIR (:95-113)
        y = y + upfirdn2d(skip, fir_up, up=2, down=1, pad=(2, 1))
    return y


def attention(x_norm, p_code, p):
    """:883-901.  groups=4, planes=out_dim/4=128, scale = planes**-0.5."""
    lr = 0.01
    n, l, _ = x_norm.shape
    m = p_code.shape[1]
    planes = p["q_transform.weight"].shape[0]
    g, gp = 4, planes // 4
    q = equal_linear(p_code, p["q_transform.weight"], p["q_transform.bias"], lr)
    k = equal_linear(x_norm, p["k_transform.weight"], p["k_transform.bias"], lr)
    v = equal_linear(x_norm, p["v_transform.weight"], p["v_transform.bias"], lr)
    q = q.reshape(n, m, g, gp).permute(0, 2, 3, 1)
    k = k.reshape(n, l, g, gp).permute(0, 2, 3, 1)
    v = v.reshape(n, l, g, gp).permute(0, 2, 3, 1)
    qk = torch.einsum("abcd,abce->abde", q, k) * planes ** -0.5
    sim = F.softmax(qk, dim=3)
    sv = torch.einsum("abcd,abed->abec", sim, v)  # N, g, gp, M
    stacked = sv.reshape(n, planes, l).permute(0, 2, 1)  # :894 (relies on M == L)
    return equal_linear(stacked, p["proj.weight"], p["proj.bias"], lr)


def attention_block(x, p_code, p):
    """:920-936.  Joint (tokens x channels) LayerNorm without affine."""
    lr = 0.01
    a = attention(F.layer_norm(x, x.shape[1:]), p_code, sub(p, "atten"))
    if "proj.weight" in p:
        x = equal_linear(x, p["proj.weight"], p["proj.bias"], lr) + a
    else:
        x = x + a
    h = F.layer_norm(x, x.shape[1:])
    h = equal_linear(h, p["mlp.0.weight"], p["mlp.0.bias"], lr)
    h = F.gelu(h)
    h = equal_linear(h, p["mlp.2.weight"], p["mlp.2.bias"], lr)
    return x + h


def interaction_stack(x0, p0, p, blocks):
    """The whole interaction network (`self.interact[i](x, spatialcode)` loop, model_spatial_query.py:668-679) over
    `attention_block` above, for parameter tables in the layout of te_attn_stack (include/te_b200.h: w_q, b_q, ...
    per block): the checker of te_attn_stack_fwd / te_attn_stack_bwd.  Block 0 reads (x0, p0), later blocks the
    previous output and p."""
    names = {"w_proj": "proj.weight", "b_proj": "proj.bias", "w_q": "atten.q_transform.weight",
             "b_q": "atten.q_transform.bias", "w_k": "atten.k_transform.weight", "b_k": "atten.k_transform.bias",
             "w_v": "atten.v_transform.weight", "b_v": "atten.v_transform.bias", "w_o": "atten.proj.weight",
             "b_o": "atten.proj.bias", "w_m1": "mlp.0.weight", "b_m1": "mlp.0.bias", "w_m2": "mlp.2.weight",
             "b_m2": "mlp.2.bias"}
    x = x0
    for i, blk in enumerate(blocks):
        params = {names[k]: v for k, v in blk.items() if k in names and v is not None}
        x = attention_block(x, p0 if i == 0 else p, params)
    return x


def map_codes(code, sd, prefix, norm_dim):
    """:626-646.  Pixel norm then one independent 512->512 linear per column."""
    code = pixel_norm(code, norm_dim)
    cols = []
    for i in range(code.shape[2]):
        cols.append(equal_linear(code[:, :, i], sd[f"{prefix}.{i + 1}.weight"],
                                 sd[f"{prefix}.{i + 1}.bias"], 0.01, activation=True))
    return torch.stack(cols, dim=2)


def generator_front(sd, z, p, n_trans=8, pixel_norm_op_dim=1, use_spatial_mapping=True,
                    use_style_mapping=True):
    """Mapping + transformer + adjust_style: (z, p) -> (latent [B,T,512], p+ [B,512,16])."""
    p_plus = map_codes(p, sd, "spatial_mapping_network", pixel_norm_op_dim) if use_spatial_mapping else p
    z_plus = map_codes(z, sd, "style_mapping_network", pixel_norm_op_dim) if use_style_mapping else z
    zt = z_plus.permute(0, 2, 1)
    pt = p_plus.permute(0, 2, 1)
    b = zt.shape[0]
    eye = sd["token_spatial"].repeat(b, 1, 1)
    x = attention_block(torch.cat([zt, eye], 2), torch.cat([pt, eye], 2), sub(sd, "interact.0"))
    for i in range(1, n_trans):
        x = attention_block(x, pt, sub(sd, f"interact.{i}"))
    latent = equal_linear(x.permute(0, 2, 1), sd["adjust_style.weight"], sd["adjust_style.bias"])
    return latent.permute(0, 2, 1), p_plus


def generator_synthesis(sd, latent, p_plus, size, noise=None, inject_noise=False):
    """:696-716.  p+ viewed [B,512,4,4] is the 4x4 input of the conv stack."""
    b = p_plus.shape[0]
    log_size = int(math.log2(size))
    fir_up = fir_kernel(gain=4.0)
    if noise is None:
        noise = [None] * ((log_size - 2) * 2 + 1)
    out = p_plus.reshape(b, 512, 4, 4)
    out = styled_conv(out, latent[:, 0], sub(sd, "conv1"), False, fir_up, noise[0], inject_noise)
    skip = to_rgb(out, latent[:, 1], sub(sd, "to_rgb1"), None, fir_up)
    li = 1
    for r in range(log_size - 2):
        out = styled_conv(out, latent[:, li], sub(sd, f"convs.{2 * r}"), True, fir_up,
                          noise[1 + 2 * r], inject_noise)
        out = styled_conv(out, latent[:, li + 1], sub(sd, f"convs.{2 * r + 1}"), False, fir_up,
                          noise[2 + 2 * r], inject_noise)
        skip = to_rgb(out, latent[:, li + 2], sub(sd, f"to_rgbs.{r}"), skip, fir_up)
        li += 2
    return skip


def generator_forward(sd, z, p, size, n_trans=8, pixel_norm_op_dim=1, noise=None,
                      inject_noise=False, use_spatial_mapping=True, use_style_mapping=True):
    """Generator.forward default route (:591-728): returns (image, latent)."""
    latent, p_plus = generator_front(sd, z, p, n_trans, pixel_norm_op_dim,
                                     use_spatial_mapping, use_style_mapping)
    img = generator_synthesis(sd, latent, p_plus, size, noise, inject_noise)
    return img, latent


# --------------------------------------------------------------------------- discriminator
def equal_conv(x, weight, bias, stride, padding):
    """:176-185"""
    cout, cin, k, _ = weight.shape
    return F.conv2d(x, weight / math.sqrt(cin * k * k), bias, stride=stride, padding=padding)


def res_block(x, p, fir):
    """:780-798"""
    y = fused_leaky_relu(equal_conv(x, p["conv1.0.weight"], None, 1, 1), p["conv1.1.bias"])
    y = upfirdn2d(y, fir, pad=(2, 2))  # Blur for k=3 downsample, :744-750
    y = fused_leaky_relu(equal_conv(y, p["conv2.1.weight"], None, 2, 0), p["conv2.2.bias"])
    s = upfirdn2d(x, fir, pad=(1, 1))  # Blur for k=1 downsample
    s = equal_conv(s, p["skip.1.weight"], None, 2, 0)
    return (y + s) / SQRT2


def discriminator_forward(sd, img):
    """:841-859"""
    fir = fir_kernel()
    x = fused_leaky_relu(equal_conv(img, sd["convs.0.0.weight"], None, 1, 0), sd["convs.0.1.bias"])
    i = 1
    while f"convs.{i}.conv1.0.weight" in sd:
        x = res_block(x, sub(sd, f"convs.{i}"), fir)
        i += 1
    b, c, h, w = x.shape
    group = min(b, 4)
    sdv = x.view(group, -1, 1, c, h, w)
    sdv = torch.sqrt(sdv.var(0, unbiased=False) + 1e-8)
    sdv = sdv.mean([2, 3, 4], keepdim=True).squeeze(2)
    sdv = sdv.repeat(group, 1, h, w)
    x = torch.cat([x, sdv], 1)
    x = fused_leaky_relu(equal_conv(x, sd["final_conv.0.weight"], None, 1, 1), sd["final_conv.1.bias"])
    x = x.view(b, -1)
    x = equal_linear(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation=True)
    return equal_linear(x, sd["final_linear.1.weight"], sd["final_linear.1.bias"])


# --------------------------------------------------------------------------- losses
def d_logistic_loss(real_pred, fake_pred):
    """train_spatial_query.py:69-73"""
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):
    """train_spatial_query.py:86-89"""
    return F.softplus(-fake_pred).mean()


def d_r1_penalty(real_pred, real_img):
    """train_spatial_query.py:76-83"""
    (g,) = torch.autograd.grad(real_pred.sum(), real_img, create_graph=True)
    return g.pow(2).reshape(g.shape[0], -1).sum(1).mean()


def g_path_lengths(fake_img, latents, noise):
    """train_spatial_query.py:92-98 with the random projection passed in explicitly."""
    (g,) = torch.autograd.grad((fake_img * noise).sum(), latents, create_graph=True)
    return torch.sqrt(g.pow(2).sum(2).mean(1))


# --------------------------------------------------------------------------- synthetic weights
def _det_randn(key, shape, salt=0):
    import zlib

    gen = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * salt) % (2 ** 31))
    return torch.randn(*shape, generator=gen)


def generator_shapes(size, channel_multiplier=2, n_trans=8):
    """State-dict layout of Generator(size, 512, 512, T) — SURVEY.md App. B.3."""
    ch = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
          128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
          512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
    log_size = int(math.log2(size))
    t = 2 * log_size - 2
    shapes = {}
    for net in ("spatial_mapping_network", "style_mapping_network"):
        for i in range(1, 17):
            shapes[f"{net}.{i}.weight"] = (512, 512)
            shapes[f"{net}.{i}.bias"] = (512,)
    shapes["adjust_style.weight"] = (t, 16)
    shapes["adjust_style.bias"] = (t,)

    def styled(prefix, cin, cout, up):
        shapes[f"{prefix}.conv.weight"] = (1, cout, cin, 3, 3)
        if up:
            shapes[f"{prefix}.conv.blur.kernel"] = (4, 4)
        shapes[f"{prefix}.conv.modulation.weight"] = (cin, 512)
        shapes[f"{prefix}.conv.modulation.bias"] = (cin,)
        shapes[f"{prefix}.noise.weight"] = (1,)
        shapes[f"{prefix}.activate.bias"] = (cout,)

    def rgb(prefix, cin, up):
        shapes[f"{prefix}.bias"] = (1, 3, 1, 1)
        if up:
            shapes[f"{prefix}.upsample.kernel"] = (4, 4)
        shapes[f"{prefix}.conv.weight"] = (1, 3, cin, 1, 1)
        shapes[f"{prefix}.conv.modulation.weight"] = (cin, 512)
        shapes[f"{prefix}.conv.modulation.bias"] = (cin,)

    styled("conv1", 512, 512, False)
    rgb("to_rgb1", 512, False)
    cin = 512
    for r, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        styled(f"convs.{2 * r}", cin, cout, True)
        styled(f"convs.{2 * r + 1}", cout, cout, False)
        rgb(f"to_rgbs.{r}", cout, True)
        cin = cout
    for li in range((log_size - 2) * 2 + 1):
        res = (li + 5) // 2
        shapes[f"noises.noise_{li}"] = (1, 1, 2 ** res, 2 ** res)
    shapes["token"] = (t, t)
    shapes["token_spatial"] = (16, 16)
    for i in range(n_trans):
        d_in = 528 if i == 0 else 512
        for nm in ("q", "k", "v"):
            shapes[f"interact.{i}.atten.{nm}_transform.weight"] = (128, d_in)
            shapes[f"interact.{i}.atten.{nm}_transform.bias"] = (128,)
        shapes[f"interact.{i}.atten.proj.weight"] = (512, 128)
        shapes[f"interact.{i}.atten.proj.bias"] = (512,)
        for j in (0, 2):
            shapes[f"interact.{i}.mlp.{j}.weight"] = (512, 512)
            shapes[f"interact.{i}.mlp.{j}.bias"] = (512,)
        if i == 0:
            shapes["interact.0.proj.weight"] = (512, 528)
            shapes["interact.0.proj.bias"] = (512,)
    return shapes


def discriminator_shapes(size, channel_multiplier=2):
    ch = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
          128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
          512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
    log_size = int(math.log2(size))
    shapes = {"convs.0.0.weight": (ch[size], 3, 1, 1), "convs.0.1.bias": (ch[size],)}
    cin = ch[size]
    for n, i in enumerate(range(log_size, 2, -1), start=1):
        cout = ch[2 ** (i - 1)]
        shapes[f"convs.{n}.conv1.0.weight"] = (cin, cin, 3, 3)
        shapes[f"convs.{n}.conv1.1.bias"] = (cin,)
        shapes[f"convs.{n}.conv2.0.kernel"] = (4, 4)
        shapes[f"convs.{n}.conv2.1.weight"] = (cout, cin, 3, 3)
        shapes[f"convs.{n}.conv2.2.bias"] = (cout,)
        shapes[f"convs.{n}.skip.0.kernel"] = (4, 4)
        shapes[f"convs.{n}.skip.1.weight"] = (cout, cin, 1, 1)
        cin = cout
    shapes["final_conv.0.weight"] = (512, 513, 3, 3)
    shapes["final_conv.1.bias"] = (512,)
    shapes["final_linear.0.weight"] = (512, 8192)
    shapes["final_linear.0.bias"] = (512,)
    shapes["final_linear.1.weight"] = (1, 512)
    shapes["final_linear.1.bias"] = (1,)
    return shapes


def synthetic_state(shapes, salt=0):
    """Deterministic, construction-order independent weights: every tensor is drawn from
    its own generator seeded by crc32(key).  Follows the reference's init statistics
    (randn weights, mapping/attention weights divided by lr_mul=0.01, :200) but makes
    biases and noise strengths non-zero so those code paths are exercised."""
    sd = {}
    for key, shape in shapes.items():
        if key.endswith("kernel"):
            # G's conv.blur / to_rgb.upsample carry FIR*4, D's blurs FIR*1 (:100,143-144)
            gain = 4.0 if key.endswith("blur.kernel") or key.endswith("upsample.kernel") else 1.0
            sd[key] = fir_kernel(gain=gain)
        elif key == "token" or key == "token_spatial":
            sd[key] = torch.eye(shape[0])
        elif key.startswith("noises."):
            sd[key] = _det_randn(key, shape, salt)
        elif key.endswith("weight"):
            w = _det_randn(key, shape, salt)
            if "mapping_network" in key or key.startswith("interact"):
                w = w / 0.01
            if key.endswith("noise.weight"):
                w = 0.1 * w
            sd[key] = w
        elif key.endswith("modulation.bias"):
            sd[key] = 1.0 + 0.1 * _det_randn(key, shape, salt)
        elif "mapping_network" in key or key.startswith("interact"):
            sd[key] = 10.0 * _det_randn(key, shape, salt)  # multiplied by lr_mul=0.01 in use
        else:
            sd[key] = 0.1 * _det_randn(key, shape, salt)
    return sd
