"""CPU emulation of the libte_b200 entry points, written from the contract in
include/te_b200.h with plain torch ops.  TEST INFRASTRUCTURE ONLY (used by the
`-m "not gpu"` host-logic tests through the `cpu_emulation` fixture)."""
import torch
import torch.nn.functional as F

from transeditor_b200 import lib


def _store(out, flat):
    """Write logical-flat storage order values into `out` (which may be channels_last)."""
    if out.dim() == 4 and not out.is_contiguous():
        n, c, h, w = out.shape
        out.copy_(flat.reshape(n, h, w, c).permute(0, 3, 1, 2).to(out.dtype))
    else:
        out.copy_(flat.reshape(out.shape).to(out.dtype))


def _storage_flat(t):
    if t.dim() == 4 and not t.is_contiguous():
        return t.permute(0, 2, 3, 1).reshape(-1)
    return t.reshape(-1)


def fused_bias_act(out, x, bias, ref, act, grad, alpha, scale, step_b, size_b):
    xs = _storage_flat(x)
    rs = _storage_flat(ref) if ref is not None else None
    flat = xs.to(torch.float64)
    if bias is not None:
        idx = (torch.arange(flat.numel()) // step_b) % size_b
        flat = flat + bias.to(torch.float64)[idx]
    r = rs.to(torch.float64) if rs is not None else torch.zeros_like(flat)
    code = act * 10 + grad
    if code in (12, 32):
        y = torch.zeros_like(flat)
    elif code == 30:
        y = torch.where(flat > 0, flat, flat * alpha)
    elif code == 31:
        y = torch.where(r > 0, flat, flat * alpha)
    else:
        y = flat
    _store(out, y * scale)


def fused_bias_act_bwd(grad_in, grad_bias, g, ref, alpha, scale, step_b, size_b):
    gs = _storage_flat(g).to(torch.float64)
    rs = _storage_flat(ref).to(torch.float64)
    y = torch.where(rs > 0, gs, gs * alpha) * scale
    _store(grad_in, y)
    if grad_bias is not None:
        idx = (torch.arange(y.numel()) // step_b) % size_b
        grad_bias.index_add_(0, idx, y.to(grad_bias.dtype))


def upfirdn2d(out, x, fir, major, in_h, in_w, minor, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    from oracle.ops_cpu import upfirdn2d_planes
    xs = _storage_flat(x).reshape(major, in_h, in_w, minor).permute(0, 3, 1, 2).reshape(major * minor, in_h, in_w)
    y = upfirdn2d_planes(xs.to(torch.float64), fir.to(torch.float64), up_x, up_y, down_x, down_y, px0, px1, py0, py1)
    oh, ow = y.shape[1], y.shape[2]
    y = y.reshape(major, minor, oh, ow).permute(0, 2, 3, 1).reshape(-1)
    _store(out, y)


def _logical_weight(w, g):
    """[cout, cin, kh, kw] view of the stored weight per te_conv_geom (w_so, w_si, flip)."""
    flat = w.reshape(-1)
    kk = g.kh * g.kw
    o = torch.arange(g.cout).view(-1, 1, 1)
    i = torch.arange(g.cin).view(1, -1, 1)
    t = torch.arange(kk).view(1, 1, -1)
    tap = (kk - 1 - t) if g.flip else t
    idx = o * g.w_so + i * g.w_si + tap
    return flat[idx].reshape(g.cout, g.cin, g.kh, g.kw), idx


def _gather_conv(x, wl, g):
    b, c, h, w = x.shape
    xz = x.new_zeros(b, c, (h - 1) * g.up + 1, (w - 1) * g.up + 1)
    xz[:, :, ::g.up, ::g.up] = x
    need_h = (g.hout - 1) * g.down + g.kh
    need_w = (g.wout - 1) * g.down + g.kw
    pr_y = need_h - g.pad_y - xz.shape[2]
    pr_x = need_w - g.pad_x - xz.shape[3]
    xp = F.pad(xz, [g.pad_x, pr_x, g.pad_y, pr_y])  # negative values crop
    return F.conv2d(xp, wl, stride=g.down)


def conv2d_simt(y, x, w, in_scale, out_scale, bias, noise, noise_w, geom):
    g = geom
    wl, _ = _logical_weight(w, g)
    xin = x if in_scale is None else x * in_scale.view(g.batch, g.cin, 1, 1)
    v = _gather_conv(xin, wl, g)
    if out_scale is not None:
        v = v * out_scale.view(g.batch, g.cout, 1, 1)
    if noise is not None:
        nz = noise.reshape(-1, 1, g.hout, g.wout)
        v = v + noise_w.reshape(()) * nz
    if bias is not None:
        v = v + bias.view(1, -1, 1, 1)
    if g.act == 1:
        v = F.leaky_relu(v, 0.2) * (2 ** 0.5)
    y.copy_(v)


def conv2d_wgrad_simt(gw, x, gy, in_scale, out_scale, geom):
    g = geom
    wl, idx = _logical_weight(torch.zeros_like(gw), g)
    wl = wl.clone().requires_grad_(True)
    xin = x if in_scale is None else x * in_scale.view(g.batch, g.cin, 1, 1)
    gyy = gy if out_scale is None else gy * out_scale.view(g.batch, g.cout, 1, 1)
    with torch.enable_grad():
        v = _gather_conv(xin.detach(), wl, g)
        (gl,) = torch.autograd.grad(v, wl, gyy.detach())
    gw.reshape(-1).index_add_(0, idx.reshape(-1), gl.reshape(-1))


def attn_core(out, sim, q, k, v, batch, tokens):
    from transeditor_b200.op import attn_core_reference
    o, s = attn_core_reference(q, k, v)
    out.copy_(o)
    if sim is not None:
        sim.copy_(s)


def adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, step, ema_decay, grad_scale):
    gr = g * grad_scale
    m.mul_(beta1).add_(gr, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)) + eps, value=-lr / bc1)
    if ema is not None:
        ema.mul_(ema_decay).add_(p, alpha=1 - ema_decay)


def scale_bc(y, x, s, batch, pixels, channels):
    xs = _storage_flat(x).reshape(batch, pixels, channels).to(torch.float64)
    _store(y, (xs * s.reshape(batch, 1, channels).to(torch.float64)).reshape(-1))


def dot_bc(out, a, b, batch, pixels, channels):
    aa = _storage_flat(a).reshape(batch, pixels, channels).to(torch.float64)
    bb = _storage_flat(b).reshape(batch, pixels, channels).to(torch.float64)
    out.add_((aa * bb).sum(1).to(out.dtype))


def adam_ema_devstep(p, g, m, v, ema, lr, beta1, beta2, eps, step_dev, ema_decay, grad_scale):
    adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, int(step_dev.item()), ema_decay, grad_scale)


def attn_stack_fwd(y, x0, p0, p, blocks, batch, lr_mul, tf32, save):
    from oracle.te_oracle import interaction_stack
    with torch.no_grad():
        y.copy_(interaction_stack(x0, p0, p, blocks))


def attn_stack_bwd(g_x0, g_p0, g_p, grads, gy, x0, p0, p, blocks, batch, lr_mul, tf32, save, gws):
    from oracle.te_oracle import interaction_stack
    fields = [f for f in lib.ATTN_FIELDS]
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)  # noqa: E731
    lx0, lp0, lp = leaf(x0), leaf(p0), leaf(p)
    lblocks = [{**{f: leaf(b.get(f)) for f in fields}, "in_dim": b["in_dim"], "param_dim": b["param_dim"]}
               for b in blocks]
    with torch.enable_grad():
        y = interaction_stack(lx0, lp0, lp, lblocks)
        wanted = [(lx0, g_x0), (lp0, g_p0)] + ([(lp, g_p)] if (lp is not None and g_p is not None) else [])
        for lb, gb in zip(lblocks, grads):
            wanted += [(lb[f], gb[f]) for f in fields if lb[f] is not None]
        got = torch.autograd.grad(y, [w for w, _ in wanted], gy, allow_unused=True)
    for (_, out), g in zip(wanted, got):
        if out is not None:
            out.copy_(torch.zeros_like(out) if g is None else g)


def _tc_gather(x, d, dy, dx):
    """x [B, C, Hin, Win] -> the anchors' input pixels for one tap, [B, C, grid_h, grid_w], zeros outside."""
    iy = torch.arange(d.grid_h) * d.in_stride + dy
    ix = torch.arange(d.grid_w) * d.in_stride + dx
    vy, vx = (iy >= 0) & (iy < d.hin), (ix >= 0) & (ix < d.win)
    g = x[:, :, iy.clamp(0, d.hin - 1)][:, :, :, ix.clamp(0, d.win - 1)]
    return g * (vy[:, None] & vx[None, :]).to(g.dtype)


def _tc_out_index(d):
    oy = torch.arange(d.grid_h) * d.out_stride + d.out_off_y
    ox = torch.arange(d.grid_w) * d.out_stride + d.out_off_x
    return oy, ox


_PAIRS = {1: [(0, 0)], 2: [(0, 0), (0, 1), (1, 0)], 3: [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]}


def _act_planes(x, nseg):
    """Activation operand -> list of f32 [B, C, H, W] planes (plain bf16 tensor, or split planes [S, B, H, W, C])."""
    if nseg == 1:
        assert x.is_contiguous(memory_format=torch.channels_last) or x.is_contiguous(), "kernel reads dense NHWC"
        return [x.float()]
    assert x.dim() == 5 and x.shape[0] == nseg and x.dtype == torch.bfloat16 and x.is_contiguous()
    return [x[i].permute(0, 3, 1, 2).float() for i in range(nseg)]


def split_bf16(dst, x, s, batch, pixels, channels, nseg):
    v = _storage_flat(x).reshape(batch, pixels, channels).float()
    if s is not None:
        v = v * s.reshape(batch, 1, channels).float()
    for i in range(nseg):
        h = v.to(torch.bfloat16)
        dst[i].reshape(batch, pixels, channels).copy_(h)
        v = v - h.float()


def conv_tc(y, x, w, out_scale, bias, desc):
    """te_conv_tc from its header contract: tap table over an anchor grid, f32 accumulation of bf16 operands (summed
    over the split-operand plane pairs), epilogue out_scale / bias / activation / residual, bf16 (or f32) output
    written at the anchors' output positions only."""
    d = desc
    nseg = d.split if d.split >= 2 else 1
    xs = _act_planes(x, nseg)
    assert w.is_contiguous(), "the kernel addresses the packed weights as a dense array"
    per_sample = d.w_bstride != 0
    wshape = (d.batch, d.w_slices, d.cout, d.cin) if per_sample else (d.w_slices, d.cout, d.cin)
    ws = [w.float().reshape(wshape)] if nseg == 1 else [w[i].float().reshape(wshape) for i in range(nseg)]
    acc = torch.zeros(d.batch, d.cout, d.grid_h, d.grid_w)
    for t in range(d.ntaps):
        for pa, pw in _PAIRS[nseg]:
            g = _tc_gather(xs[pa], d, d.tap_dy[t], d.tap_dx[t])
            if per_sample:
                acc += torch.einsum("bchw,boc->bohw", g, ws[pw][:, d.tap_w[t]])
            else:
                acc += torch.einsum("bchw,oc->bohw", g, ws[pw][d.tap_w[t]])
    if out_scale is not None:
        acc = acc * out_scale.reshape(d.batch, d.cout, 1, 1)
    if bias is not None:
        acc = acc + bias.reshape(1, d.cout, 1, 1)
    residual, slope = getattr(d, "py_refs", (None, None))
    if d.act == 3:
        acc = torch.where(acc < 0, acc * slope.reshape(1, -1, 1, 1), acc)
    elif d.act != 0:
        gain = d.act_gain if d.act_gain != 0 else (1.0 if d.act == 2 else 2 ** 0.5)
        acc = F.leaky_relu(acc, 0.01 if d.act == 2 else 0.2) * gain
    oy, ox = _tc_out_index(d)
    if residual is not None:
        acc = acc + residual.float()[:, :, oy][:, :, :, ox]
    y[:, :, oy[:, None], ox[None, :]] = acc.to(y.dtype)


def conv_wgrad_tc(gw, g, x, desc):
    d = desc
    alpha = d.wgrad_alpha if d.wgrad_alpha != 0 else 1.0
    oy, ox = _tc_out_index(d)
    nseg = d.split if d.split >= 2 else 1
    gas = [gp[:, :, oy][:, :, :, ox] for gp in _act_planes(g, nseg)]   # gradient at the anchors' output positions
    xps = _act_planes(x, nseg)
    per_sample = d.w_bstride != 0
    view = gw.view((d.batch, d.w_slices, d.cout, d.cin) if per_sample else (d.w_slices, d.cout, d.cin))
    for t in range(d.ntaps):
        for pg, px in _PAIRS[nseg]:
            xs = _tc_gather(xps[px], d, d.tap_dy[t], d.tap_dx[t])
            if per_sample:
                view[:, d.tap_w[t]] += alpha * torch.einsum("bohw,bchw->boc", gas[pg], xs)
            else:
                view[d.tap_w[t]] += alpha * torch.einsum("bohw,bchw->oc", gas[pg], xs)


def pack_weights_tc(tasks):
    for task in tasks:
        src, dst_n, dst_t, scale = task[:4]
        nseg = task[4] if len(task) > 4 else 1
        o, i, k, _ = src.shape
        v = (src.detach() * scale).permute(2, 3, 0, 1).reshape(k * k, o, i)  # f32 product, like the kernel
        for sg in range(nseg):
            h = v.to(torch.bfloat16)
            if dst_n is not None:
                (dst_n if nseg == 1 else dst_n[sg])[:, :o, :i].copy_(h)
            if dst_t is not None:
                (dst_t if nseg == 1 else dst_t[sg])[:, :i, :o].copy_(h.transpose(1, 2))
            v = v - h.float()


def linear_grouped(tasks, tf32=False):
    for t in tasks:
        x, w = t["x"].double(), t["w"].double()
        if t.get("pixel_norm"):
            r = torch.rsqrt((x * x).mean(1, keepdim=True) + 1e-8)
            if t.get("rnorm_out") is not None:
                t["rnorm_out"].copy_(r[:, 0].float())
            x = x * r
        y = (x @ (w if t.get("w_trans") else w.t())) * t.get("alpha", 1.0)
        if t.get("bias") is not None:
            y = y + t["bias"].double() * t.get("bias_mul", 1.0)
        if t.get("act"):
            y = F.leaky_relu(y, 0.2) * 2 ** 0.5
        if t.get("k_splits", 1) > 1:
            t["y"].add_(y.float())
        else:
            t["y"].copy_(y.float())


def linear_wgrad_grouped(tasks):
    for t in tasks:
        x = t["x"].double()
        if t.get("x_scale") is not None:
            x = x * t["x_scale"].double().unsqueeze(1)
        g = t["g"].double()
        t["gw"].copy_(((g.t() @ x) * t.get("alpha", 1.0)).float())
        if t.get("gbias") is not None:
            t["gbias"].copy_((g.sum(0) * t.get("bias_mul", 1.0)).float())


def from_rgb_fwd(y, x, w, bias, wscale, slope, gain):
    u = F.conv2d(x.double(), (w.double() * wscale).view(w.shape[0], 3, 1, 1))
    if bias is not None:
        u = u + bias.double().view(1, -1, 1, 1)
    y.copy_((F.leaky_relu(u, slope) * gain).to(y.dtype))


def from_rgb_bwd(gw, gbias, gx, g, out, x, w, wscale, slope, gain):
    gp = g.double() * torch.where(out.double() > 0, gain, gain * slope)
    gw.add_((torch.einsum("bohw,bchw->oc", gp, x.double()) * wscale).float())
    if gbias is not None:
        gbias.add_(gp.sum((0, 2, 3)).float())
    if gx is not None:
        gx.copy_((torch.einsum("bohw,oc->bchw", gp, w.double()) * wscale).float())


def wgrad_unpack(out, ws, batch, o_dim, i_dim, taps, rows, ld, trans, clear=False):
    w = ws.reshape(batch, taps, rows, ld)
    src = w[:, :, :i_dim, :o_dim].permute(0, 3, 2, 1) if trans else w[:, :, :o_dim, :i_dim].permute(0, 2, 3, 1)
    out.copy_(src.reshape(out.shape))
    if clear:
        (w[:, :, :i_dim, :o_dim] if trans else w[:, :, :o_dim, :i_dim]).zero_()


def upfirdn2d_bias_act(out, x, fir, bias, major, in_h, in_w, minor, px0, px1, py0, py1, slope, gain):
    return False   # the emulation always takes the two-call route (upfirdn2d, then fused_bias_act)


def weight_energy(energy, w, rows, taps, coef):
    energy.copy_((w.reshape(rows, taps).double().square().sum(1) * coef).reshape(energy.shape).float())


def weight_energy_bwd(gw, w, g, rows, taps, coef2):
    gw.copy_((w.reshape(rows, taps).double() * g.reshape(rows, 1).double() * coef2).reshape(gw.shape).float())


def image_prep(dst_nchw, dst_nhwc8, src_hwc, flip, batch, h, w):
    x = src_hwc
    if flip is not None:
        x = torch.stack([img.flip(1) if int(f) else img for img, f in zip(x, flip)])
    t = x.to(torch.float32).div(255).sub(0.5).div(0.5)
    if dst_nchw is not None:
        dst_nchw.copy_(t.permute(0, 3, 1, 2))
    if dst_nhwc8 is not None:
        dst_nhwc8.zero_()
        dst_nhwc8[..., :3] = t.to(dst_nhwc8.dtype)


def image_quantize(dst_hwc, src, low, high):
    t = src.float().clamp(low, high).sub(low).div(max(high - low, 1e-5))
    dst_hwc.copy_(t.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8))


def install(monkeypatch):
    monkeypatch.setattr(lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(lib, "emulated", True)
    for name in ("fused_bias_act", "fused_bias_act_bwd", "upfirdn2d", "conv2d_simt",
                 "conv2d_wgrad_simt", "attn_core", "adam_ema", "adam_ema_devstep", "scale_bc", "dot_bc",
                 "attn_stack_fwd", "attn_stack_bwd", "pack_weights_tc", "conv_tc", "conv_wgrad_tc", "split_bf16",
                 "image_prep", "image_quantize", "linear_grouped", "linear_wgrad_grouped", "from_rgb_fwd",
                 "from_rgb_bwd", "wgrad_unpack", "upfirdn2d_bias_act", "weight_energy",
                 "weight_energy_bwd"):
        monkeypatch.setattr(lib, name, globals()[name])
