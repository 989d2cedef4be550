"""Operator layer: the reference's `utils.op` surface (fused_leaky_relu, FusedLeakyReLU,
upfirdn2d) plus the convolution primitives, all as twice-differentiable autograd
Functions over the C ABI of libte_b200.so.

Reference interfaces mirrored (argument names, defaults, error behaviour):
  utils/op/fused_act.py:18-90   FusedLeakyReLUFunction(+Backward), FusedLeakyReLU, fused_leaky_relu
  utils/op/upfirdn2d.py:17-148  UpFirDn2d(+Backward), upfirdn2d

Every backward is itself expressed through differentiable Functions of this file, so
R1 (double backward through D) and path-length regularisation (double backward through
G) work — SURVEY.md fact 5.
"""
import math

import torch
from torch import nn
from torch.autograd import Function

from . import lib

# --------------------------------------------------------------------------------------------
# layout helpers


def _is_channels_last(t):
    return (t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


def _canonical(t):
    """Return (tensor, channels_last_flag) with the storage dense in one of the two layouts."""
    if _is_channels_last(t):
        return t, True
    return t.contiguous(), False


def _bias_geometry(x, channels_last):
    """(step_b, size_b) of te_fused_bias_act: bias indexes dim 1 of x."""
    c = x.shape[1] if x.dim() > 1 else x.shape[0]
    if channels_last or x.dim() <= 2:
        return 1, c
    return int(math.prod(x.shape[2:])), c


# --------------------------------------------------------------------------------------------
# fused bias + leaky relu


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    lib.require_cuda(x, bias, ref)
    x, cl = _canonical(x)
    if ref is not None:
        ref = ref.contiguous(memory_format=torch.channels_last) if cl else ref.contiguous()
    if bias is not None:
        # 16-bit activations take the f32 master bias directly (TE_BIAS_F32): no down-cast kernel per call
        if not (bias.dtype == torch.float32 and x.dtype in (torch.bfloat16, torch.float16)):
            bias = bias.to(x.dtype)
        bias = bias.contiguous()
    out = torch.empty_like(x)
    step_b, size_b = _bias_geometry(x, cl)
    lib.fused_bias_act(out, x, bias, ref, act, grad, float(alpha), float(scale), step_b, size_b)
    return out


class FusedLeakyReLUFunctionBackward(Function):
    """grad_input = scale * (out > 0 ? g : slope*g); grad_bias = per-channel sum — ONE kernel
    (the reference runs fused_bias_act and then a separate .sum(), fused_act.py:27-36)."""

    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale, bias_dtype=None):
        lib.require_cuda(grad_output, out)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        ctx.has_bias = has_bias
        g, cl = _canonical(grad_output)
        ref = out.contiguous(memory_format=torch.channels_last) if cl else out.contiguous()
        grad_input = torch.empty_like(g)
        step_b, size_b = _bias_geometry(g, cl)
        if has_bias:
            acc_dtype = torch.float64 if g.dtype == torch.float64 else torch.float32
            grad_bias = torch.zeros(size_b, dtype=acc_dtype, device=g.device)
            lib.fused_bias_act_bwd(grad_input, grad_bias, g, ref, float(negative_slope), float(scale),
                                   step_b, size_b)
            grad_bias = grad_bias.to(bias_dtype or g.dtype)
        else:
            lib.fused_bias_act_bwd(grad_input, None, g, ref, float(negative_slope), float(scale),
                                   step_b, size_b)
            grad_bias = grad_input.new_zeros(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        if gradgrad_input is None:
            gradgrad_input = torch.zeros_like(out)
        gb = gradgrad_bias if (ctx.has_bias and gradgrad_bias is not None) else None
        gradgrad_out = _bias_act(gradgrad_input, gb, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None, None


class FusedLeakyReLUFunction(Function):
    """Saves only its OUTPUT, like the reference (fused_act.py:52-59)."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        ctx.has_bias = bias is not None
        ctx.bias_dtype = bias.dtype if bias is not None else None
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(
            grad_output, out, ctx.has_bias, ctx.negative_slope, ctx.scale, ctx.bias_dtype)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """utils/op/fused_act.py:89-90."""
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """utils/op/fused_act.py:72-86 — parameter name `bias`, shape [channel]."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


# --------------------------------------------------------------------------------------------
# upfirdn2d


def _fir_f32(kernel):
    return kernel.detach().to(torch.float32).contiguous()


def _upfirdn2d_native(x, fir, up, down, pad, out_hw=None):
    """x: [N,C,H,W] in either dense layout -> [N,C,H',W'] in the same layout."""
    lib.require_cuda(x, fir)
    x, cl = _canonical(x)
    n, c, h, w = x.shape
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    kh, kw = fir.shape
    out_h = (h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (w * up_x + px0 + px1 - kw) // down_x + 1
    if out_hw is not None and (out_h, out_w) != tuple(out_hw):
        raise RuntimeError("upfirdn2d: size mismatch %s vs %s" % ((out_h, out_w), tuple(out_hw)))
    if out_h <= 0 or out_w <= 0:
        raise RuntimeError("upfirdn2d: empty output")
    if cl:
        out = torch.empty((n, c, out_h, out_w), dtype=x.dtype, device=x.device,
                          memory_format=torch.channels_last)
        major, minor = n, c
    else:
        out = torch.empty((n, c, out_h, out_w), dtype=x.dtype, device=x.device)
        major, minor = n * c, 1
    lib.upfirdn2d(out, x, _fir_f32(fir), major, h, w, minor, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
    return out


class UpFirDn2dBackward(Function):
    """utils/op/upfirdn2d.py:17-83: the input gradient is the same operator with up and down
    swapped, the flipped FIR and g_pad; its own backward is the forward operator again."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        grad_input = _upfirdn2d_native(grad_output, grad_kernel, down, up, g_pad,
                                       out_hw=(in_size[2], in_size[3]))
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.out_size = out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        (kernel,) = ctx.saved_tensors
        gradgrad_out = _upfirdn2d_native(gradgrad_input, kernel, ctx.up, ctx.down, ctx.pad,
                                         out_hw=ctx.out_size)
        return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """utils/op/upfirdn2d.py:86-140."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        _, _, in_h, in_w = input.shape
        ctx.in_size = tuple(input.shape)
        out = _upfirdn2d_native(input, kernel, up, down, pad)
        out_h, out_w = out.shape[2], out.shape[3]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        # padding of the gradient operator (upfirdn2d.py:109-112)
        ctx.g_pad = (kernel_w - pad_x0 - 1,
                     in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1,
                     in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down,
                                             ctx.pad, ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


class BlurBiasAct(Function):
    """fused_leaky_relu(upfirdn2d(x, kernel, pad=pad), bias, slope, gain) for channels-last activations in ONE pass
    (te_upfirdn2d_bias_act: the blur's output never round-trips through HBM before the activation); two launches when
    the geometry is not the fused kernel's.  Backward is the composition of the differentiable pieces — the masked
    gradient from FusedLeakyReLUFunctionBackward (sign of the saved OUTPUT), then the blur's adjoint — so second order
    works exactly as for the unfused ops."""

    @staticmethod
    def forward(ctx, x, kernel, pad, bias, negative_slope, scale):
        lib.require_cuda(x, kernel, bias)
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        xc, cl = _canonical(x)
        n, c, h, w = xc.shape
        out = None
        if cl and 0 <= negative_slope <= 1 and scale > 0:
            oh, ow = h + py0 + py1 - kh + 1, w + px0 + px1 - kw + 1
            out = torch.empty((n, c, oh, ow), dtype=xc.dtype, device=xc.device, memory_format=torch.channels_last)
            if not lib.upfirdn2d_bias_act(out, xc, _fir_f32(kernel), bias.detach().float().contiguous(), n, h, w, c, px0,
                                          px1, py0, py1, float(negative_slope), float(scale)):
                out = None
        if out is None:
            out = _bias_act(_upfirdn2d_native(xc, kernel, (1, 1), (1, 1), pad), bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]), out)
        ctx.in_size, ctx.out_size, ctx.pad = tuple(x.shape), (out.shape[2], out.shape[3]), pad
        ctx.g_pad = (kw - px0 - 1, w - out.shape[3] + px0, kh - py0 - 1, h - out.shape[2] + py0)
        ctx.negative_slope, ctx.scale, ctx.bias_dtype = negative_slope, scale, bias.dtype
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel, out = ctx.saved_tensors
        g_v, g_b = FusedLeakyReLUFunctionBackward.apply(grad_output, out, True, ctx.negative_slope, ctx.scale,
                                                        ctx.bias_dtype)
        g_x = None
        if ctx.needs_input_grad[0]:
            g_x = UpFirDn2dBackward.apply(g_v, kernel, grad_kernel, (1, 1), (1, 1), ctx.pad, ctx.g_pad, ctx.in_size,
                                          ctx.out_size)
        return g_x, None, None, (g_b if ctx.needs_input_grad[3] else None), None, None


def blur_bias_act(input, kernel, pad, bias, negative_slope=0.2, scale=2 ** 0.5):
    """upfirdn2d(input, kernel, pad=pad) + bias -> leaky ReLU * scale, fused (same pad on both axes)."""
    return BlurBiasAct.apply(input, kernel, (pad[0], pad[1], pad[0], pad[1]), bias, negative_slope, scale)


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """utils/op/upfirdn2d.py:143-148 — same pad on both axes."""
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


# --------------------------------------------------------------------------------------------
# gather convolution primitives (f32 / f64, NCHW) — see include/te_b200.h


class Geometry:
    """up/down/pad/flip of one gather convolution; weights are STORED [O, I, kh, kw] relative to
    the forward op and `transposed` says the kernel should read them as [I, O]."""

    __slots__ = ("kh", "kw", "up", "down", "pad_y", "pad_x", "flip", "transposed", "out_hw")

    def __init__(self, kh, kw, up, down, pad_y, pad_x, flip, transposed, out_hw):
        self.kh, self.kw, self.up, self.down = kh, kw, up, down
        self.pad_y, self.pad_x, self.flip, self.transposed = pad_y, pad_x, flip, transposed
        self.out_hw = (int(out_hw[0]), int(out_hw[1]))

    def adjoint(self, in_hw):
        """Geometry of the data gradient (maps output-shaped tensors back to `in_hw`)."""
        return Geometry(self.kh, self.kw, self.down, self.up, self.kh - 1 - self.pad_y,
                        self.kw - 1 - self.pad_x, not self.flip, not self.transposed, in_hw)

    def c_struct(self, x, w, act=0, noise_bstride=0):
        b, cin, hin, win = x.shape
        o_store, i_store = w.shape[0], w.shape[1]
        kk = self.kh * self.kw
        if self.transposed:
            cout, w_so, w_si, w_cin = i_store, kk, i_store * kk, o_store
        else:
            cout, w_so, w_si, w_cin = o_store, i_store * kk, kk, i_store
        if w_cin != cin:
            raise RuntimeError("conv: weight expects %d input channels, tensor has %d" % (w_cin, cin))
        return lib.ConvGeom(b, cin, hin, win, cout, self.out_hw[0], self.out_hw[1], self.kh, self.kw,
                            self.up, self.down, self.pad_y, self.pad_x, int(self.flip),
                            w_so, w_si, act, noise_bstride)


def _conv_forward(x, w, geom, in_scale=None, out_scale=None, bias=None, noise=None, noise_w=None, act=0):
    lib.require_cuda(x, w, in_scale, out_scale, bias, noise, noise_w)
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("conv2d_simt: f32/f64 only, got %s" % x.dtype)
    x = x.contiguous()
    w = w.to(x.dtype).contiguous()
    nb = 0
    if noise is not None:
        noise = noise.to(x.dtype).contiguous()
        nb = 0 if noise.shape[0] == 1 else geom.out_hw[0] * geom.out_hw[1]
        noise_w = noise_w.to(x.dtype).contiguous()
    cs = geom.c_struct(x, w, act, nb)
    y = torch.empty((cs.batch, cs.cout, cs.hout, cs.wout), dtype=x.dtype, device=x.device)
    prep = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    lib.conv2d_simt(y, x, w, prep(in_scale), prep(out_scale), prep(bias), noise, noise_w, cs)
    return y


def _conv_wgrad(x, gy, geom, w_shape, in_scale=None, out_scale=None):
    lib.require_cuda(x, gy)
    x = x.contiguous()
    gy = gy.to(x.dtype).contiguous()
    gw = torch.zeros(w_shape, dtype=x.dtype, device=x.device)
    cs = geom.c_struct(x, gw)
    if (cs.hout, cs.wout) != (gy.shape[2], gy.shape[3]) or cs.cout != gy.shape[1]:
        raise RuntimeError("conv wgrad: gradient shape %s does not match geometry" % (tuple(gy.shape),))
    prep = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    lib.conv2d_wgrad_simt(gw, x, gy, prep(in_scale), prep(out_scale), cs)
    return gw


class ConvGather(Function):
    """y = conv(x, w; geom).  backward: data grad = ConvGather with the adjoint geometry,
    weight grad = ConvWeightGrad — both differentiable again."""

    @staticmethod
    def forward(ctx, x, w, geom):
        ctx.save_for_backward(x, w)
        ctx.geom = geom
        return _conv_forward(x, w, geom)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        geom = ctx.geom
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = ConvGather.apply(gy, w, geom.adjoint((x.shape[2], x.shape[3])))
        if ctx.needs_input_grad[1]:
            gw = ConvWeightGrad.apply(x, gy, geom, tuple(w.shape))
        return gx, gw, None


class ConvWeightGrad(Function):
    """gw = wgrad(x, gy; geom) — bilinear, so its backward is two ConvGather calls."""

    @staticmethod
    def forward(ctx, x, gy, geom, w_shape):
        ctx.save_for_backward(x, gy)
        ctx.geom = geom
        return _conv_wgrad(x, gy, geom, w_shape)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        geom = ctx.geom
        g_x = g_gy = None
        if ctx.needs_input_grad[0]:
            g_x = ConvGather.apply(gy, ggw, geom.adjoint((x.shape[2], x.shape[3])))
        if ctx.needs_input_grad[1]:
            g_gy = ConvGather.apply(x, ggw, geom)
        return g_x, g_gy, None, None


def conv2d(x, w, stride=1, padding=0):
    """F.conv2d(x, w, stride=stride, padding=padding) (cross-correlation), w [O, I, kh, kw]."""
    kh, kw = w.shape[2], w.shape[3]
    oh = (x.shape[2] + 2 * padding - kh) // stride + 1
    ow = (x.shape[3] + 2 * padding - kw) // stride + 1
    geom = Geometry(kh, kw, 1, stride, padding, padding, False, False, (oh, ow))
    return ConvGather.apply(x, w, geom)


def conv_transpose2d(x, w_oi, stride=2):
    """F.conv_transpose2d(x, w_oi.transpose(0,1), stride=stride, padding=0) with the weight kept in
    the modulated-conv storage order [O, I, kh, kw] (model_spatial_query.py:315-318)."""
    kh, kw = w_oi.shape[2], w_oi.shape[3]
    oh = (x.shape[2] - 1) * stride + kh
    ow = (x.shape[3] - 1) * stride + kw
    geom = Geometry(kh, kw, stride, 1, kh - 1, kw - 1, True, False, (oh, ow))
    return ConvGather.apply(x, w_oi, geom)


def conv2d_fused(x, w, in_scale=None, out_scale=None, bias=None, noise=None, noise_w=None,
                 act=False, stride=1, padding=0, transpose_stride=0):
    """Inference-only fused form: act(out_scale * conv(x * in_scale, w) + noise_w*noise + bias).
    Not differentiable (callers use it under no_grad)."""
    kh, kw = w.shape[2], w.shape[3]
    if transpose_stride:
        oh = (x.shape[2] - 1) * transpose_stride + kh
        ow = (x.shape[3] - 1) * transpose_stride + kw
        geom = Geometry(kh, kw, transpose_stride, 1, kh - 1, kw - 1, True, False, (oh, ow))
    else:
        oh = (x.shape[2] + 2 * padding - kh) // stride + 1
        ow = (x.shape[3] + 2 * padding - kw) // stride + 1
        geom = Geometry(kh, kw, 1, stride, padding, padding, False, False, (oh, ow))
    return _conv_forward(x, w, geom, in_scale, out_scale, bias, noise, noise_w, 1 if act else 0)


# --------------------------------------------------------------------------------------------
# per-(sample, channel) scale and its reduction on channels-last tensors (modulation / demodulation)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


class ScaleBC(Function):
    """y[b,c,h,w] = x[b,c,h,w] * s[b,c]  (x channels-last, s f32).  Differentiable to any order
    together with DotBC."""

    @staticmethod
    def forward(ctx, x, s):
        lib.require_cuda(x, s)
        x = _cl(x)
        b, c, h, w = x.shape
        sf = s.to(torch.float32).contiguous()
        y = torch.empty_like(x, memory_format=torch.channels_last)
        lib.scale_bc(y, x, sf, b, h * w, c)
        ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, s = ctx.saved_tensors
        gx = ScaleBC.apply(gy, s) if ctx.needs_input_grad[0] else None
        gs = DotBC.apply(gy, x).to(s.dtype) if ctx.needs_input_grad[1] else None
        return gx, gs


class DotBC(Function):
    """out[b,c] = sum_hw a * b  (f32 accumulate, f32 result)."""

    @staticmethod
    def forward(ctx, a, b_):
        lib.require_cuda(a, b_)
        a, b_ = _cl(a), _cl(b_.to(a.dtype))
        n, c, h, w = a.shape
        out = torch.zeros((n, c), dtype=torch.float32, device=a.device)
        lib.dot_bc(out, a, b_, n, h * w, c)
        ctx.save_for_backward(a, b_)
        return out

    @staticmethod
    def backward(ctx, go):
        a, b_ = ctx.saved_tensors
        ga = ScaleBC.apply(b_, go) if ctx.needs_input_grad[0] else None
        gb = ScaleBC.apply(a, go) if ctx.needs_input_grad[1] else None
        return ga, gb


def scale_bc(x, s):
    return ScaleBC.apply(x, s)


# --------------------------------------------------------------------------------------------
# cross-attention core


class AttnCore(Function):
    """q,k,v [B,16,128] -> (out [B,16,128], sim [B,4,16,16]) in one kernel
    (model_spatial_query.py:888-894).  Backward recomputes with differentiable torch ops so
    higher-order gradients (spatial path regularisation) stay available."""

    @staticmethod
    def forward(ctx, q, k, v):
        lib.require_cuda(q, k, v)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        b, t, c = q.shape
        out = torch.empty_like(q)
        sim = torch.empty((b, 4, t, t), dtype=q.dtype, device=q.device)
        lib.attn_core(out, sim, q, k, v, b, t)
        ctx.save_for_backward(q, k, v)
        ctx.mark_non_differentiable(sim)
        return out, sim

    @staticmethod
    def backward(ctx, g_out, _g_sim):
        # closed-form softmax-attention gradient written with differentiable torch ops
        q, k, v = ctx.saved_tensors
        b, t, c = q.shape
        h, d = 4, c // 4
        scale = c ** -0.5
        split = lambda z: z.reshape(b, t, h, d).permute(0, 2, 1, 3)  # noqa: E731
        merge = lambda z: z.permute(0, 2, 1, 3).reshape(b, t, c)  # noqa: E731
        qh, kh, vh, go = split(q), split(k), split(v), split(g_out)
        sim = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
        g_v = sim.transpose(-1, -2) @ go
        g_sim = go @ vh.transpose(-1, -2)
        g_logit = sim * (g_sim - (g_sim * sim).sum(-1, keepdim=True)) * scale
        g_q = g_logit @ kh
        g_k = g_logit.transpose(-1, -2) @ qh
        return merge(g_q), merge(g_k), merge(g_v)


def attn_core_reference(q, k, v):
    """Same math with torch ops (used for the recompute backward only)."""
    b, t, c = q.shape
    h, d = 4, c // 4
    qh = q.reshape(b, t, h, d).permute(0, 2, 1, 3)
    kh = k.reshape(b, t, h, d).permute(0, 2, 1, 3)
    vh = v.reshape(b, t, h, d).permute(0, 2, 1, 3)
    sim = torch.softmax(qh @ kh.transpose(-1, -2) * (c ** -0.5), dim=-1)
    out = (sim @ vh).permute(0, 2, 1, 3).reshape(b, t, c)
    return out, sim


def attn_core(q, k, v):
    if q.dtype != torch.float32 or q.shape[1] != 16 or q.shape[2] != 128:
        return attn_core_reference(q, k, v)  # f64 gradcheck / unusual token counts
    return AttnCore.apply(q, k, v)


# --------------------------------------------------------------------------------------------
# the whole interaction network (all AttentionBlocks) in one launch

_ATTN_FIELDS = lib.ATTN_FIELDS


def _equal_linear(x, w, b, lr_mul):
    """EqualLinear (model_spatial_query.py:194-221) as one addmm."""
    x2 = x.reshape(-1, x.shape[-1])
    out = torch.addmm(b, x2, w.t(), beta=lr_mul, alpha=lr_mul / math.sqrt(w.shape[1]))
    return out.reshape(*x.shape[:-1], w.shape[0])


def attn_stack_reference(x0, p0, p, blocks, lr_mul):
    """The stack written with differentiable ops (`AttentionBlock.forward` :920-936 over `Attention.forward`
    :883-901): what te_attn_stack_fwd computes.  Used when a gradient OF the gradient is requested (the optional
    spatial path regulariser, train_spatial_query.py:252-277), for dtypes the kernel does not take, and as the
    definition the tests compare the kernel with."""
    x = x0
    for i, w in enumerate(blocks):
        pm = p0 if i == 0 else p
        xn = torch.nn.functional.layer_norm(x, x.shape[1:])
        q = _equal_linear(pm, w["w_q"], w["b_q"], lr_mul)
        k = _equal_linear(xn, w["w_k"], w["b_k"], lr_mul)
        v = _equal_linear(xn, w["w_v"], w["b_v"], lr_mul)
        att, _ = attn_core(q, k, v)
        a = _equal_linear(att, w["w_o"], w["b_o"], lr_mul)
        x = (_equal_linear(x, w["w_proj"], w["b_proj"], lr_mul) if w.get("w_proj") is not None else x) + a
        h = _equal_linear(torch.nn.functional.layer_norm(x, x.shape[1:]), w["w_m1"], w["b_m1"], lr_mul)
        x = x + _equal_linear(torch.nn.functional.gelu(h), w["w_m2"], w["b_m2"], lr_mul)
    return x


def _blocks_from_flat(flat, dims):
    n = len(_ATTN_FIELDS)
    blocks = []
    for i, (din, dp) in enumerate(dims):
        blk = dict(zip(_ATTN_FIELDS, flat[i * n:(i + 1) * n]))
        blk["in_dim"], blk["param_dim"] = din, dp
        blocks.append(blk)
    return blocks


class AttnStack(Function):
    """y = interact(x0, p0, p) through te_attn_stack_fwd; first-order backward through te_attn_stack_bwd (two
    launches).  Under create_graph=True the backward is re-expressed with `attn_stack_reference`, so second-order
    gradients exist (they are only needed by the optional spatial path regulariser)."""

    @staticmethod
    def forward(ctx, x0, p0, p, lr_mul, dims, tf32, *flat):
        lib.require_cuda(x0, p0, p, *flat)
        blocks = _blocks_from_flat(flat, dims)
        batch = x0.shape[0]
        y = torch.empty((batch, 16, 512), dtype=torch.float32, device=x0.device)
        need = any(ctx.needs_input_grad)
        save = None
        if need:
            n_save, _ = lib.attn_stack_workspace(blocks, batch)
            save = torch.empty(n_save, dtype=torch.float32, device=x0.device)
        lib.attn_stack_fwd(y, x0, p0, p, blocks, batch, lr_mul, tf32, save)
        ctx.lr_mul, ctx.dims, ctx.has_p, ctx.tf32 = lr_mul, dims, p is not None, tf32
        ctx.save_for_backward(x0, p0, p if p is not None else x0.new_empty(0), save if need else x0.new_empty(0), *flat)
        return y

    @staticmethod
    def backward(ctx, gy):
        x0, p0, p, save, *flat = ctx.saved_tensors
        p = p if ctx.has_p else None
        dims, lr_mul = ctx.dims, ctx.lr_mul
        if torch.is_grad_enabled():  # create_graph=True: stay differentiable
            with torch.enable_grad():
                # PARTIAL derivatives with respect to each argument are wanted, but the arguments may depend on each
                # other in the caller's graph (Generator.forward builds p0 = cat(p, eye) from p): differentiate with
                # respect to fresh view nodes, which no other argument descends from, and which still lead back to
                # the originals for the second-order terms
                alias = lambda t: None if t is None else t.view_as(t)  # noqa: E731
                ax0, ap0, ap = alias(x0), alias(p0), alias(p)
                aflat = [alias(t) for t in flat]
                live = [t for t in (ax0, ap0, ap, *aflat) if t is not None and t.requires_grad]
                y = attn_stack_reference(ax0, ap0, ap, _blocks_from_flat(aflat, dims), lr_mul)
                got = dict(zip(map(id, live), torch.autograd.grad(y, live, gy, create_graph=True, allow_unused=True)))
            pick = lambda t: None if t is None else got.get(id(t))  # noqa: E731
            return (pick(ax0), pick(ap0), pick(ap), None, None, None, *[pick(t) for t in aflat])
        blocks = _blocks_from_flat(flat, dims)
        batch = x0.shape[0]
        _, n_gws = lib.attn_stack_workspace(blocks, batch)
        gws = torch.empty(n_gws, dtype=torch.float32, device=x0.device)
        g_x0, g_p0 = torch.empty_like(x0), torch.empty_like(p0)
        g_p = torch.empty_like(p) if (p is not None and len(dims) > 1) else None
        # one buffer for all parameter gradients; the kernel writes every element
        sizes = [0 if t is None else t.numel() for t in flat]
        gflat = torch.empty(sum(sizes), dtype=torch.float32, device=x0.device)
        views, off = [], 0
        for t, n in zip(flat, sizes):
            views.append(None if t is None else gflat[off:off + n].view(t.shape))
            off += n
        lib.attn_stack_bwd(g_x0, g_p0, g_p, _blocks_from_flat(views, dims), gy.contiguous(), x0, p0, p, blocks, batch,
                           lr_mul, ctx.tf32, save, gws)
        if p is not None and g_p is None:
            g_p = torch.zeros_like(p)
        return (g_x0, g_p0, g_p, None, None, None, *views)


def attn_stack_supported(x0, p0, p, blocks):
    """Shapes / dtypes te_attn_stack_fwd is built for (include/te_b200.h)."""
    if x0.dtype != torch.float32 or x0.dim() != 3 or x0.shape[1] != 16 or p0.shape[1] != 16:
        return False
    if not 1 <= len(blocks) <= 8:
        return False
    for i, w in enumerate(blocks):
        din, dp = w["in_dim"], w["param_dim"]
        if din % 16 or dp % 16 or not (64 <= din <= 528 and 64 <= dp <= 528):
            return False
        if any(w[f] is not None and w[f].data_ptr() % 16 for f in _ATTN_FIELDS if f.startswith("w_")):
            return False
        if i > 0 and (din != 512 or dp != 512):
            return False
        if (w.get("w_proj") is not None) != (din != 512):
            return False
        if tuple(w["w_q"].shape) != (128, dp) or tuple(w["w_o"].shape) != (512, 128) or tuple(w["w_m1"].shape) != (512, 512):
            return False
    return x0.shape[2] == blocks[0]["in_dim"] and p0.shape[2] == blocks[0]["param_dim"] and \
        (p is None or tuple(p.shape[1:]) == (16, 512))


def attn_stack(x0, p0, p, blocks, lr_mul, tf32=False):
    """blocks: list of dicts {w_proj, b_proj (None unless in_dim != 512), w_q, b_q, w_k, b_k, w_v, b_v, w_o, b_o,
    w_m1, b_m1, w_m2, b_m2, in_dim, param_dim}.  tf32: single-pass TF32 products (the bf16 mode's choice, like
    allow_tf32 for the library GEMMs) instead of the fp32-equivalent 3xTF32.  Returns [B, 16, 512]."""
    flat = [w.get(f) for w in blocks for f in _ATTN_FIELDS]
    dims = tuple((w["in_dim"], w["param_dim"]) for w in blocks)
    # made contiguous HERE (differentiably), so that the tensors the Function saves are its own inputs
    return AttnStack.apply(x0.contiguous(), p0.contiguous(), None if p is None else p.contiguous(), float(lr_mul),
                           dims, bool(tf32), *flat)


# --------------------------------------------------------------------------------------------
# tensor-core (tcgen05) convolution, bf16 channels-last


def pack_weight_tc(w):
    """[O, I, kh, kw] (any float dtype) -> tap-major bf16 [kh*kw, O, I] (K contiguous) for te_conv2d_tc.
    A leading batch dim ([B, O, I, kh, kw] -> [B, kh*kw, O, I]) gives per-sample weights."""
    if w.dim() == 5:
        b, o, i, kh, kw = w.shape
        return w.permute(0, 3, 4, 1, 2).reshape(b, kh * kw, o, i).to(torch.bfloat16).contiguous()
    o, i, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, o, i).to(torch.bfloat16).contiguous()


def conv2d_tc(x, w_packed, ksize, out_scale=None, bias=None, act=False):
    """x: [B, Cin, H, W] bf16 with channels-last strides; stride 1, padding ksize//2.
    Returns [B, Cout, H, W] bf16 channels-last.  Not differentiable by itself."""
    lib.require_cuda(x, w_packed, out_scale, bias)
    if x.dtype != torch.bfloat16:
        raise TypeError("conv2d_tc: bf16 activations required, got %s" % x.dtype)
    b, cin, h, w = x.shape
    x = x.contiguous(memory_format=torch.channels_last)
    per_sample = w_packed.dim() == 4
    cout = w_packed.shape[-2]
    if w_packed.shape[-1] != cin or w_packed.shape[-3] != ksize * ksize:
        raise RuntimeError("conv2d_tc: weight %s does not match Cin=%d k=%d" % (tuple(w_packed.shape), cin, ksize))
    y = torch.empty((b, cout, h, w), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()  # noqa: E731
    lib.conv2d_tc(y, x, w_packed, f32(out_scale), f32(bias), b, h, w, cin, cout, ksize, ksize,
                  1 if act else 0, ksize * ksize * cout * cin if per_sample else 0)
    return y


# --------------------------------------------------------------------------------------------
# Grouped EqualLinear (te_linear_grouped / te_linear_wgrad_grouped): the small-M linears of the path, many per launch
_LRELU_GAIN = 2 ** 0.5


def _linear_composite(x, w, b, alpha, bias_mul, act, pixel_norm):
    """Differentiable restatement of one task with torch ops (model_spatial_query.py:75-81, 194-221): used when a
    gradient of the gradient is asked for (create_graph=True)."""
    if pixel_norm:
        x = x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)
    y = (x @ w.t()) * alpha
    if b is not None:
        y = y + b * bias_mul
    if act:
        y = torch.nn.functional.leaky_relu(y, 0.2) * _LRELU_GAIN
    return y


class GroupedLinear(Function):
    """ys[i] = act_i(alpha_i * x_i @ w_i^T + b_i * bias_mul_i) for a LIST of independent EqualLinear layers in one
    launch; the backward pass is one grouped data-gradient launch plus one grouped weight/bias-gradient launch.
    meta: tuple of (alpha, bias_mul, act, k_splits) per task; tensors: x_0, w_0, b_0, x_1, w_1, b_1, ... (b_i may be
    None).  k_splits > 1 cuts a long reduction (D's 8192-wide final linear) over several CTAs; no activation then."""

    @staticmethod
    def forward(ctx, meta, tf32, *tensors):
        n = len(meta)
        xs, ws, bs = tensors[0::3], tensors[1::3], tensors[2::3]
        lib.require_cuda(*[t for t in tensors if t is not None])
        total = sum(x.shape[0] * w.shape[0] for x, w in zip(xs, ws))
        split = any(ks > 1 for _, _, _, ks in meta)   # split-K tasks accumulate into zeros
        flat = (torch.zeros if split else torch.empty)(total, dtype=torch.float32, device=xs[0].device)
        ys, tasks, off = [], [], 0
        for (alpha, bias_mul, act, ks), x, w, b in zip(meta, xs, ws, bs):
            m, nn_ = x.shape[0], w.shape[0]
            y = flat[off:off + m * nn_].view(m, nn_)
            off += m * nn_
            ys.append(y)
            tasks.append(dict(x=x, w=w.contiguous(), y=y, bias=b, bias_mul=bias_mul, alpha=alpha, act=int(act),
                              k_splits=ks))
        lib.linear_grouped(tasks, tf32)
        ctx.meta, ctx.tf32, ctx.n = meta, tf32, n
        ctx.save_for_backward(*[t if t is not None else flat.new_empty(0) for t in tensors], *ys)
        ctx.has_bias = [b is not None for b in bs]
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gys):
        n, meta = ctx.n, ctx.meta
        saved = ctx.saved_tensors
        tensors, ys = saved[:3 * n], saved[3 * n:]
        xs, ws = tensors[0::3], tensors[1::3]
        bs = [b if hb else None for b, hb in zip(tensors[2::3], ctx.has_bias)]
        grads = [None] * (3 * n)
        if torch.is_grad_enabled():  # create_graph=True: stay differentiable (see AttnStack.backward)
            with torch.enable_grad():
                for i, ((alpha, bias_mul, act, _), x, w, b, gy) in enumerate(zip(meta, xs, ws, bs, gys)):
                    if gy is None:
                        continue
                    alias = [t.view_as(t) if t is not None else None for t in (x, w, b)]
                    live = [t for t in alias if t is not None and t.requires_grad]
                    if not live:
                        continue
                    y = _linear_composite(alias[0], alias[1], alias[2], alpha, bias_mul, act, False)
                    got = dict(zip(map(id, live), torch.autograd.grad(y, live, gy, create_graph=True, allow_unused=True)))
                    for j, t in enumerate(alias):
                        grads[3 * i + j] = None if t is None else got.get(id(t))
            return (None, None, *grads)
        dtasks, wtasks = [], []
        for i, ((alpha, bias_mul, act, _), x, w, b, y, gy) in enumerate(zip(meta, xs, ws, bs, ys, gys)):
            if gy is None:
                continue
            g = gy
            if act:  # d leaky_relu(u) * gain: the sign of the output is the sign of u (gain > 0)
                g = gy * torch.where(y > 0, _LRELU_GAIN, 0.2 * _LRELU_GAIN)
            if ctx.needs_input_grad[2 + 3 * i]:
                gx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
                dtasks.append(dict(x=g, w=w.contiguous(), w_trans=True, y=gx, alpha=alpha))
                grads[3 * i] = gx
            need_w = ctx.needs_input_grad[3 + 3 * i]
            need_b = b is not None and ctx.needs_input_grad[4 + 3 * i]
            if need_w or need_b:
                gw = torch.empty(w.shape, dtype=torch.float32, device=x.device)
                gb = torch.empty(b.shape, dtype=torch.float32, device=x.device) if need_b else None
                wtasks.append(dict(g=g, x=x, gw=gw, gbias=gb, alpha=alpha, bias_mul=bias_mul))
                grads[3 * i + 1] = gw if need_w else None
                grads[3 * i + 2] = gb
        lib.linear_grouped(dtasks, ctx.tf32)
        lib.linear_wgrad_grouped(wtasks)
        return (None, None, *grads)


def grouped_linear(layers, tf32=False):
    """layers: list of (x [M,K], weight [N,K], bias [N] or None, alpha, bias_mul, act[, k_splits]) -> list of y [M,N]."""
    meta = tuple((float(l[3]), float(l[4]), bool(l[5]), int(l[6]) if len(l) > 6 else 1) for l in layers)
    flat = []
    for l in layers:
        flat += [l[0], l[1], l[2]]
    return list(GroupedLinear.apply(meta, tf32, *flat))


class MappingColumns(Function):
    """The pre-mapping of one code (model_spatial_query.py:626-646): PixelNorm over dim 1 of code [B, D, C], then column
    i through its own EqualLinear(D, D, lr_mul, 'fused_lrelu') — `count` layers, ONE launch, the normalisation, bias
    and activation inside it.  Returns [B, D, C] (columns >= count stay zero, :630).  weights / biases: the layers'
    parameters in column order."""

    @staticmethod
    def forward(ctx, code, alpha, lr_mul, tf32, count, pixel_norm, *params):
        lib.require_cuda(code, *params)
        ws, bs = params[:count], params[count:]
        b, d, c = code.shape
        cols = code.permute(2, 0, 1).contiguous()                       # [C, B, D]: unit-stride rows for the kernel
        out = torch.empty_like(code) if count == c else torch.zeros_like(code)
        rn = torch.empty((count, b) if pixel_norm else (0, b), dtype=torch.float32, device=code.device)
        tasks = [dict(x=cols[i], w=ws[i].contiguous(), bias=bs[i], bias_mul=lr_mul, y=out[:, :, i], alpha=alpha, act=1,
                      pixel_norm=pixel_norm, rnorm_out=rn[i] if pixel_norm else None) for i in range(count)]
        lib.linear_grouped(tasks, tf32)
        ctx.alpha, ctx.lr_mul, ctx.tf32, ctx.count, ctx.pixel_norm = alpha, lr_mul, tf32, count, pixel_norm
        ctx.save_for_backward(code, cols, rn, out, *params)
        return out

    @staticmethod
    def backward(ctx, gout):
        code, cols, rn, out, *params = ctx.saved_tensors
        count, alpha, lr_mul = ctx.count, ctx.alpha, ctx.lr_mul
        ws, bs = params[:count], params[count:]
        if torch.is_grad_enabled():  # create_graph=True
            with torch.enable_grad():
                acode = code.view_as(code)
                aparams = [t.view_as(t) for t in params]
                live = [t for t in (acode, *aparams) if t.requires_grad]
                y = mapping_columns_reference(acode, aparams[:count], aparams[count:], alpha, lr_mul, ctx.pixel_norm)
                got = dict(zip(map(id, live), torch.autograd.grad(y, live, gout, create_graph=True, allow_unused=True)))
            return (got.get(id(acode)), None, None, None, None, None, *[got.get(id(t)) for t in aparams])
        b, d, c = code.shape
        g = gout * torch.where(out > 0, _LRELU_GAIN, 0.2 * _LRELU_GAIN)      # [B, D, C]
        gcols = g.permute(2, 0, 1).contiguous()                                # [C, B, D]
        gw = torch.empty((count, d, d), dtype=torch.float32, device=code.device)
        gb = torch.empty((count, d), dtype=torch.float32, device=code.device)
        pn = ctx.pixel_norm
        lib.linear_wgrad_grouped([dict(g=gcols[i], x=cols[i], x_scale=rn[i] if pn else None, gw=gw[i], gbias=gb[i],
                                       alpha=alpha, bias_mul=lr_mul) for i in range(count)])
        gcode = None
        if ctx.needs_input_grad[0]:
            gxn = torch.empty((count, b, d), dtype=torch.float32, device=code.device)
            lib.linear_grouped([dict(x=gcols[i], w=ws[i].contiguous(), w_trans=True, y=gxn[i], alpha=alpha)
                                for i in range(count)], ctx.tf32)
            # PixelNorm backward: xn = x * r, r = (mean x^2 + eps)^-1/2  =>  gx = r * gxn - x * r^3 * mean(gxn * x)
            gx = gxn
            if pn:
                x, r = cols[:count], rn.unsqueeze(-1)
                gx = r * gxn - x * (r * r * r) * (gxn * x).mean(-1, keepdim=True)
            gcode = torch.zeros_like(code) if count != c else torch.empty_like(code)
            gcode[:, :, :count] = gx.permute(1, 2, 0)
        return (gcode, None, None, None, None, None, *gw.unbind(0), *gb.unbind(0))


def mapping_columns_reference(code, weights, biases, alpha, lr_mul, pixel_norm=True):
    """Differentiable torch restatement of MappingColumns (the literal loop of model_spatial_query.py:626-632)."""
    x = code * torch.rsqrt(torch.mean(code ** 2, dim=1, keepdim=True) + 1e-8) if pixel_norm else code
    cols = []
    for i, (w, b) in enumerate(zip(weights, biases)):
        cols.append(_linear_composite(x[:, :, i], w, b, alpha, lr_mul, True, False))
    y = torch.stack(cols, dim=2)
    if len(cols) == code.shape[2]:
        return y
    out = torch.cat([y, torch.zeros_like(code[:, :, len(cols):])], dim=2)
    return out


def mapping_columns(code, weights, biases, alpha, lr_mul, tf32=False, pixel_norm=True):
    """pixel_norm=True: PixelNorm over dim 1 (the per-column feature vector) is computed inside the kernel; pass False
    when the caller has normalised over another dimension already."""
    return MappingColumns.apply(code, float(alpha), float(lr_mul), tf32, len(weights), bool(pixel_norm), *weights,
                                *biases)


# --------------------------------------------------------------------------------------------
# The discriminator's from-RGB layer (te_from_rgb_fwd / te_from_rgb_bwd)
def from_rgb_reference(img, weight, bias, wscale, gain, slope=0.2):
    """Differentiable torch restatement: EqualConv2d(3, C, 1) + FusedLeakyReLU (model_spatial_query.py:731-777)."""
    y = torch.nn.functional.conv2d(img, weight * wscale) + bias.view(1, -1, 1, 1)
    return torch.nn.functional.leaky_relu(y, slope) * gain


class FromRGB(Function):
    """out = leaky_relu(conv1x1(img, W * wscale) + bias, 0.2) * gain, img f32 NCHW [B,3,H,W] -> channels-last
    [B,C,H,W] of `dtype`; one streaming kernel forward, one backward (weight, bias and image gradients in a single pass
    over the activation and its gradient, LeakyReLU mask applied on the fly).  create_graph=True (R1) re-expresses the
    backward with differentiable ops."""

    @staticmethod
    def forward(ctx, img, weight, bias, wscale, gain, dtype):
        lib.require_cuda(img, weight, bias)
        if img.dtype != torch.float32 or img.dim() != 4 or img.shape[1] != 3:
            raise TypeError("from_rgb: expected a float32 [B, 3, H, W] image, got %s %s" % (img.dtype, tuple(img.shape)))
        img = img.contiguous()
        b, _, h, w = img.shape
        c = weight.shape[0]
        out = torch.empty((b, c, h, w), dtype=dtype, device=img.device, memory_format=torch.channels_last)
        w2 = weight.detach().reshape(c, 3).contiguous()
        lib.from_rgb_fwd(out, img, w2, bias.detach().float().contiguous(), wscale, 0.2, gain)
        ctx.save_for_backward(img, weight, bias, out)
        ctx.wscale, ctx.gain = wscale, gain
        return out

    @staticmethod
    def backward(ctx, g_out):
        img, weight, bias, out = ctx.saved_tensors
        wscale, gain = ctx.wscale, ctx.gain
        if torch.is_grad_enabled():
            # create_graph=True (the R1 penalty): re-express the layer with the twice-differentiable tensor-core ops
            # (zero-padded to a 64-channel operand, tc.TcConvBiasAct) and differentiate that
            from . import tc
            with torch.enable_grad():
                alias = [t.view_as(t) for t in (img, weight, bias)]
                live = [t for t in alias if t.requires_grad]
                b, _, h, w = img.shape
                xcl = torch.zeros((b, h, w, 64), dtype=out.dtype, device=img.device)
                xcl = torch.cat([alias[0].permute(0, 2, 3, 1).to(out.dtype), xcl[..., 3:]], dim=3).permute(0, 3, 1, 2)
                wpad = torch.nn.functional.pad(alias[1], (0, 0, 0, 0, 0, 61))
                y = tc.conv2d_bias_act(xcl, wpad, alias[2], wscale=wscale, gain=gain)
                got = dict(zip(map(id, live), torch.autograd.grad(y, live, g_out.to(y.dtype), create_graph=True,
                                                                  allow_unused=True)))
            return (*[got.get(id(t)) for t in alias], None, None, None)
        c = weight.shape[0]
        g = g_out.to(out.dtype).contiguous(memory_format=torch.channels_last)
        acc = torch.zeros(c * 4, dtype=torch.float32, device=img.device)   # [C,3] weight gradient + [C] bias gradient
        gw, gb = acc[:c * 3].view(c, 3), acc[c * 3:]
        gx = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        lib.from_rgb_bwd(gw, gb, gx, g, out, img, weight.detach().reshape(c, 3).contiguous(), wscale, 0.2, gain)
        return gx, gw.view(weight.shape), gb.to(bias.dtype), None, None, None


def from_rgb(img, weight, bias, wscale, gain=2 ** 0.5, dtype=torch.bfloat16):
    return FromRGB.apply(img, weight, bias, float(wscale), float(gain), dtype)
