"""Generator / Discriminator of TransEditor on the B200 kernels.

Host-side mirror of /root/reference/model_spatial_query.py: same class names, constructor
and forward signatures, return conventions and `state_dict` layout, so
the reference's train_spatial_query.py / test_spatial_query.py and the authors' checkpoints
work unmodified.  What differs is everything underneath:

* ModulatedConv2d never materialises per-sample weights (reference :299-317 writes a
  [B*Cout, Cin, k, k] tensor and calls cuDNN with groups=B).  It runs the shared-weight form
  d * conv(x * s, W) on our own convolution kernels.
* All convolutions, upfirdn2d, bias+activation and the attention core are hand-written
  sm_100a kernels behind the C ABI (include/te_b200.h); there is no CPU path.
* The 2 x 16 per-column mapping linears run as one batched GEMM each instead of 32 small
  GEMMs + 32 slice copies (reference :626-646).

Citations `:N` are into the reference's model_spatial_query.py.
"""
import math
import os

import torch
from torch import nn
from torch.nn import functional as F

from . import op, tc
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d

_SQRT2 = math.sqrt(2.0)

# Precision of the convolution stacks.
#   "fp32"       the parity mode: f32 channels-last activations, every convolution on the tcgen05 kernels in
#                SPLIT-OPERAND form (tc.py: bf16 hi/mid planes, three tensor-core products per f32 product, f32
#                accumulation spread over four TMEM accumulators) — measured image max-abs error 2.2e-4 at 256^2
#                and 4.5e-4 at 1024^2 against the 1e-3 bar; tc.set_split_planes(3) halves it at twice the cost.
#   "bf16"       the speed mode: channels-last bf16 activations, plain bf16 tcgen05 products (f32 accumulation, f32
#                master weights, f32 mapping / transformer / RGB skip path).
#   "fp32_simt"  NCHW f32 on the SIMT gather kernels (exact f32 FMA chains; kept for f64 gradchecks and as a
#                cross-check of the split-operand mode).
# Layers dispatch on the dtype of their input (`_tc`); this flag decides the conversion at the model boundaries and
# which engine f32 tensors use.
_PRECISION = "fp32"


def set_precision(name):
    global _PRECISION
    if name not in ("fp32", "bf16", "fp32_simt"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'fp32_simt'")
    _PRECISION = name


def get_precision():
    return _PRECISION


_SMALL_M = 64   # EqualLinear: up to this many rows go through the grouped weight-streaming kernel


def lib_emulated():
    """True only inside the CPU test-suite's `cpu_emulation` fixture (tests/emu.py sets lib.emulated)."""
    from . import lib
    return lib.emulated


def _tc(x):
    """True when the tensor-core engine handles activations like x."""
    return x.dtype == torch.bfloat16 or (x.dtype == torch.float32 and _PRECISION == "fp32")


def _tc_mode():
    return _PRECISION in ("bf16", "fp32")


def _act_dtype():
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


def _to_cl(x, dtype=None, pad_to=8):
    """NCHW -> channels-last `dtype` (default: the precision mode's activation dtype) with the channel count
    padded to a multiple of `pad_to`."""
    dtype = dtype or _act_dtype()
    b, c, h, w = x.shape
    if c % pad_to == 0:
        return x.to(dtype).contiguous(memory_format=torch.channels_last)
    # one memset of the padded buffer + one strided convert-copy of the real channels (instead of pad in
    # f32, dtype conversion and a layout change: three passes over the PADDED tensor)
    buf = torch.zeros((b, h, w, c + pad_to - c % pad_to), dtype=dtype, device=x.device)
    buf[..., :c] = x.permute(0, 2, 3, 1)
    return buf.permute(0, 3, 1, 2)


def _to_bf16_cl(x, pad_to=8):
    return _to_cl(x, torch.bfloat16, pad_to)


class _DenseGrad(torch.autograd.Function):
    """Identity whose gradient comes back dense in the INPUT's layout.  The tensor-core routes run channels-last
    inside, so d(out)/d(image) would reach the caller as a permuted view; the reference's R1 penalty calls
    `.view(batch, -1)` on it (train_spatial_query.py:81), which needs the NCHW-contiguous gradient cuDNN returns."""

    @staticmethod
    def forward(ctx, x):
        ctx.channels_last = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.contiguous(memory_format=torch.channels_last) if ctx.channels_last else g.contiguous()


_SIDE_STREAMS = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def _grad_needed(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class PixelNorm(nn.Module):
    """:75-81"""

    def __init__(self, pixel_norm_op_dim):
        super().__init__()
        self.pixel_norm_op_dim = pixel_norm_op_dim

    def forward(self, input):
        ms = input.square().mean(dim=self.pixel_norm_op_dim, keepdim=True)
        return input * torch.rsqrt(ms + 1e-8)


def make_kernel(k):
    """:84-92 — separable FIR as a normalised outer product."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


class Upsample(nn.Module):
    """:95-113"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """:116-134 (unused by G/D, kept for API completeness)"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """:137-153"""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """:156-191 — conv2d with the equalised-lr weight scale 1/sqrt(Cin k^2)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def _tc_weight(self, input):
        """Master weight for the tensor-core route; the equalised-lr scale is applied while it is repacked."""
        k = self.weight.shape[2]
        if self.padding != (k // 2 if self.stride == 1 else 0) or self.stride not in (1, 2):
            raise RuntimeError("tensor-core EqualConv2d supports stride 1 (pad k//2) and stride 2 (pad 0)")
        w = self.weight
        if input.shape[1] != w.shape[1]:  # input channels were zero-padded
            w = F.pad(w, (0, 0, 0, 0, 0, input.shape[1] - w.shape[1]))
        return w

    def _tc_ok(self, input):
        k = self.weight.shape[2]
        return (_tc(input) and self.weight.shape[0] % 8 == 0 and self.stride in (1, 2)
                and self.padding == (k // 2 if self.stride == 1 else 0)
                and (input.dtype == torch.bfloat16 or k in (1, 3)))

    def forward(self, input):
        if self._tc_ok(input):
            if input.shape[1] % 8:  # odd channel counts (RGB): zero channels, matched by zero weight columns
                input = F.pad(input, (0, 0, 0, 0, 0, 8 - input.shape[1] % 8))
            out = tc.conv2d(input, self._tc_weight(input), stride=self.stride, wscale=self.scale)
            if self.bias is not None:
                out = out + self.bias.view(1, -1, 1, 1).to(out.dtype)
            return out
        w = self.weight * self.scale
        out = op.conv2d(input, w, stride=self.stride, padding=self.padding)
        if self.bias is not None:
            out = out + self.bias.view(1, -1, 1, 1)
        return out

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},"
                f" {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})")


class EqualLinear(nn.Module):
    """:194-226"""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1.0, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        x2 = input.reshape(-1, input.shape[-1])
        if x2.shape[0] <= _SMALL_M and x2.dtype == torch.float32 and (x2.is_cuda or lib_emulated()):
            # small batch of rows: weight-streaming bound -> te_linear_grouped (scale, bias * lr_mul and the fused
            # leaky ReLU inside the kernel; data / weight / bias gradients are two more launches)
            k = x2.shape[1]
            if k >= 2048:  # long reduction (D's 8192 -> 512): cut over several CTAs, bias / activation afterwards
                (out,) = op.grouped_linear([(x2, self.weight, None, self.scale, 1.0, False, k // 1024)],
                                           tf32=_PRECISION == "bf16")
                if self.activation:
                    out = fused_leaky_relu(out, self.bias * self.lr_mul)
                elif self.bias is not None:
                    out = out + self.bias * self.lr_mul
            else:
                (out,) = op.grouped_linear([(x2, self.weight, self.bias, self.scale, self.lr_mul,
                                             bool(self.activation))], tf32=_PRECISION == "bf16")
            return out.reshape(*input.shape[:-1], out.shape[-1])
        # many rows (adjust_style: B*512 rows of 16): a plain library GEMM.  x (W*scale)^T + b*lr_mul as ONE addmm with
        # alpha=scale: the reference's separate `weight * scale` elementwise kernel (:215,218) and its backward
        # disappear (beta=0 ignores the placeholder input)
        if self.activation or self.bias is None:
            out = torch.addmm(x2.new_empty(1), x2, self.weight.t(), beta=0, alpha=self.scale)
        else:  # beta folds the bias's lr_mul too: no `bias * lr_mul` kernel, forward or backward
            out = torch.addmm(self.bias, x2, self.weight.t(), beta=self.lr_mul, alpha=self.scale)
        out = out.reshape(*input.shape[:-1], out.shape[-1])
        if self.activation:
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ScaledLeakyReLU(nn.Module):
    """:229-238"""

    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return fused_leaky_relu(input, None, self.negative_slope, _SQRT2)


class WeightEnergy(torch.autograd.Function):
    """energy[o, i] = sum_k (W[o, i, k] * scale)^2, the weight-only factor of the demodulation coefficient
    (model_spatial_query.py:301-303 with the style factored out).  One node instead of autograd's
    mul / square / sum chain; the backward is written with differentiable ops."""

    @staticmethod
    def forward(ctx, w, scale):
        ctx.save_for_backward(w)
        ctx.scale = scale
        if (w.is_cuda or lib_emulated()) and w.dtype == torch.float32 and w.is_contiguous() \
                and os.environ.get("TE_WEIGHT_ENERGY", "1") != "0":
            from . import lib
            o, i, kh, kw = w.shape
            energy = torch.empty((o, i), dtype=torch.float32, device=w.device)
            lib.weight_energy(energy, w.detach(), o * i, kh * kw, scale * scale)   # one row of taps per thread
            return energy
        # ||W[o,i,:]||^2 * scale^2 with ONE pass over the weight (the literal mul / square / sum chain is three)
        return torch.linalg.vector_norm(w, dim=(2, 3)).square_().mul_(scale * scale)

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        if not torch.is_grad_enabled() and (w.is_cuda or lib_emulated()) and w.dtype == torch.float32 \
                and w.is_contiguous() and g.dtype == torch.float32 and os.environ.get("TE_WEIGHT_ENERGY", "1") != "0":
            from . import lib
            o, i, kh, kw = w.shape
            gw = torch.empty_like(w)
            lib.weight_energy_bwd(gw, w.detach(), g.contiguous(), o * i, kh * kw, 2.0 * ctx.scale * ctx.scale)
            return gw, None
        return w * (g * (2.0 * ctx.scale * ctx.scale))[:, :, None, None], None   # differentiable form (create_graph)


class ModulatedConv2d(nn.Module):
    """:241-337.  Shared-weight formulation: with Wn = W/sqrt(Cin k^2), s = modulation(style),
    d[b,o] = rsqrt(sum_i s[b,i]^2 * sum_k Wn[o,i,k]^2 + 1e-8):

        plain     y = d * conv(x * s, Wn, pad k//2)
        upsample  y = blur( d * conv_transpose(x * s, Wn, stride 2) )   (blur is linear, d is
                                                                         per-channel: they commute)
        downsample (unused by G) y = d * conv(blur(x * s), Wn, stride 2)
    """

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True,
                 upsample=False, downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1),
                             upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def scales(self, style, fold_scale=False, s=None):
        """(s [B,Cin], d [B,Cout] or None, Wn [Cout,Cin,k,k]); with fold_scale the raw master weight is returned
        instead of Wn = W * scale (the tensor-core route applies `scale` while repacking the weight).  `s`: the
        modulation output when the caller has computed it already (all layers in one grouped launch)."""
        if s is None:
            s = self.modulation(style)
        w = self.weight[0]
        d = None
        if self.demodulate:
            energy = WeightEnergy.apply(w, self.scale)  # [Cout, Cin] = sum_k (W * scale)^2
            d = torch.rsqrt(s.square() @ energy.t() + self.eps)
        return s, d, (w if fold_scale else w * self.scale)

    def forward(self, input, style, bias=None, noise=None, noise_weight=None, activate=False):
        """`bias`/`noise`/`activate` let StyledConv hand its epilogue to the conv kernel on the
        inference path; reference callers pass only (input, style)."""
        if _tc(input) and (input.dtype == torch.bfloat16 or (self.in_channel % 8 == 0 and not self.downsample)):
            ops = self._take_prepared(input)
            if ops is None:
                ops = self.tc_operands(style, input.shape[2] * input.shape[3])
            fused_ok = not _grad_needed(input, ops["s"], ops["w"], bias, noise_weight)
            return self._forward_tc(input, ops, bias, noise, noise_weight, activate, fused_ok)
        s, d, wn = self.scales(style)
        fused_ok = not _grad_needed(input, s, wn, bias, noise_weight)
        if self.upsample:
            if fused_ok:
                out = op.conv2d_fused(input, wn, in_scale=s, out_scale=d, transpose_stride=2)
            else:
                out = op.conv_transpose2d(input * s[:, :, None, None], wn, stride=2)
                if d is not None:
                    out = out * d[:, :, None, None]
            out = self.blur(out)
            return _epilogue(out, bias, noise, noise_weight, activate)
        if self.downsample:
            x = self.blur(input * s[:, :, None, None])
            out = op.conv2d(x, wn, stride=2, padding=0)
            if d is not None:
                out = out * d[:, :, None, None]
            return _epilogue(out, bias, noise, noise_weight, activate)
        if fused_ok:
            return op.conv2d_fused(input, wn, in_scale=s, out_scale=d, bias=bias, noise=noise,
                                   noise_w=noise_weight, act=activate, padding=self.padding)
        out = op.conv2d(input * s[:, :, None, None], wn, stride=1, padding=self.padding)
        if d is not None:
            out = out * d[:, :, None, None]
        return _epilogue(out, bias, noise, noise_weight, activate)


def _modconv_tc_operands(self, style, hw, s=None):
    """Everything of the bf16 route that depends only on (style, weights): s, d and — at high resolution —
    the per-sample weights.  No activation is touched, so Generator.forward computes these for ALL layers on a
    side stream while the main stream runs the convolutions (`prepare`)."""
    if self.downsample:
        raise RuntimeError("tensor-core ModulatedConv2d: downsample is not used by the generator")
    s, d, wn = self.scales(style, fold_scale=True, s=s)
    k = self.kernel_size
    cout = self.out_channel
    if cout % 8:  # ToRGB: pad the 3 output channels to 8 (zero rows), sliced off after the conv
        wn = F.pad(wn, (0, 0, 0, 0, 0, 0, 0, 8 - cout % 8))
        if d is not None:
            d = F.pad(d, (0, 8 - cout % 8), value=1.0)
    ops = {"s": s, "d": d, "w": wn, "wb": None}
    if hw >= 128 and hw >= 4 * wn.shape[0] * k * k:
        # High resolution: per-sample weights W*s*d (the reference's own formulation, :299-304) are tiny
        # next to the activations (1-15 % of their size), so fold modulation AND demodulation into them
        # and skip every activation-sized scaling pass, forward and backward.
        wb = wn.unsqueeze(0) * s[:, None, :, None, None]
        if d is not None:
            wb = wb * d[:, :, None, None, None]
        ops["wb"] = wb
    return ops


def _modconv_prepare(self, style, hw, side, main, s=None):
    """Compute tc_operands on the stream `side`; forward() picks them up after waiting on the event."""
    with torch.cuda.stream(side):
        ops = self.tc_operands(style, hw, s=s)
        ev = torch.cuda.Event()
        ev.record(side)
    # temporaries allocated on the side stream but consumed (and later freed) under the main stream
    temps = [ops["s"], ops["d"], ops["wb"]]
    if ops["w"].data_ptr() != self.weight.data_ptr():
        temps.append(ops["w"])
    for t in temps:
        if t is not None and t.is_cuda:
            t.record_stream(main)
    self._prepared = (ops, ev, hw)


def _modconv_take_prepared(self, x):
    prep = getattr(self, "_prepared", None)
    if prep is None:
        return None
    self._prepared = None
    ops, ev, hw = prep
    if hw != x.shape[2] * x.shape[3] or ops["s"].shape[0] != x.shape[0]:
        return None
    torch.cuda.current_stream(x.device).wait_event(ev)
    return ops


def _modconv_forward_tc(self, x, ops, bias, noise, noise_weight, activate, fused_ok):
    """bf16 channels-last route on the tcgen05 kernels: y = d * conv(x * s, W * wscale) as
    scale_bc -> conv (-> blur) -> scale_bc, every piece a twice-differentiable custom op.  Without
    autograd the demodulation, bias and activation ride in the conv kernel's epilogue.  ops["w"] is the RAW
    master weight; the equalised-lr scale is applied when the weight is repacked to bf16."""
    s, d, wn, wb = ops["s"], ops["d"], ops["w"], ops["wb"]
    wscale = self.scale
    k = self.kernel_size
    cout = self.out_channel
    if wb is not None:
        if self.upsample:
            v = tc.conv_transpose2d(x, wb, wscale=wscale)
            if activate and noise is None and bias is not None and wn.shape[0] == cout:
                # blur -> bias -> leaky ReLU in ONE pass over the activation (op.BlurBiasAct)
                return op.blur_bias_act(v, self.blur.kernel, self.blur.pad, bias.reshape(-1))
            v = self.blur(v)
        elif fused_ok and noise is None:
            pb = None if bias is None else F.pad(bias.reshape(-1), (0, wn.shape[0] - cout))
            v = tc.conv_raw(x, tc.pack_weight(wb, False, wscale, tc.nseg_for(x)), tc.Mode("s1", k), bias=pb,
                            act=activate)
            return v if wn.shape[0] == cout else v[:, :cout]
        elif activate and noise is None and bias is not None and wn.shape[0] == cout:
            return tc.conv2d_bias_act(x, wb, bias.reshape(-1), wscale=wscale)  # bias + lrelu in the epilogue
        else:
            v = tc.conv2d(x, wb, wscale=wscale)
        if wn.shape[0] != cout:
            v = v[:, :cout]
        return _epilogue(v, bias, noise, noise_weight, activate)
    # Low resolution: shared weights, modulation / demodulation applied to the (small) activations.
    if fused_ok and noise is None and not self.upsample:
        pb = None if bias is None else F.pad(bias.reshape(-1), (0, wn.shape[0] - cout))
        wp = tc.pack_weight(wn, False, wscale, tc.nseg_for(x))
        if x.dtype == torch.float32:  # the modulation rides in the operand split
            v = tc.conv_raw(x, wp, tc.Mode("s1", k), out_scale=d, bias=pb, act=activate, in_scale=s)
        else:
            v = tc.conv_raw(op.scale_bc(x, s), wp, tc.Mode("s1", k), out_scale=d, bias=pb, act=activate)
        return v if wn.shape[0] == cout else v[:, :cout]
    u = op.scale_bc(x, s)
    if d is not None and wn.shape[0] == cout and os.environ.get("TE_CONV_SCALED", "1") != "0":
        # demodulation in the convolution epilogue (tc.TcConvScaled; d commutes with the blur: both are per-channel
        # linear), and for the upsampling layers blur + bias + leaky ReLU in one pass behind it
        if self.upsample:
            v = tc.conv_transpose2d_scaled(u, wn, d, wscale=wscale)
            if activate and noise is None and bias is not None:
                return op.blur_bias_act(v, self.blur.kernel, self.blur.pad, bias.reshape(-1))
            v = self.blur(v)
        else:
            v = tc.conv2d_scaled(u, wn, d, wscale=wscale)
        return _epilogue(v, bias, noise, noise_weight, activate)
    if self.upsample:
        v = self.blur(tc.conv_transpose2d(u, wn, wscale=wscale))
    else:
        v = tc.conv2d(u, wn, wscale=wscale)
    if d is not None:
        v = op.scale_bc(v, d)
    if wn.shape[0] != cout:
        v = v[:, :cout]
    return _epilogue(v, bias, noise, noise_weight, activate)


ModulatedConv2d._forward_tc = _modconv_forward_tc
ModulatedConv2d.tc_operands = _modconv_tc_operands
ModulatedConv2d.prepare = _modconv_prepare
ModulatedConv2d._take_prepared = _modconv_take_prepared


def _epilogue(out, bias, noise, noise_weight, activate):
    if noise is not None:
        out = out + (noise_weight * noise).to(out.dtype)
    if activate:
        return fused_leaky_relu(out, bias)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1).to(out.dtype)
    return out


class NoiseInjection(nn.Module):
    """:340-351"""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """:354-364 (unused: the P code is the 4x4 input)"""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """:367-403 — modulated conv -> [noise] -> bias + leaky relu * sqrt(2)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False,
                 blur_kernel=[1, 3, 3, 1], demodulate=True, layer_noise_injection=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim,
                                    upsample=upsample, blur_kernel=blur_kernel, demodulate=demodulate)
        self.layer_noise_injection = layer_noise_injection
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        if self.layer_noise_injection:
            if noise is None:
                b, _, h, w = input.shape
                f = 2 if self.conv.upsample else 1
                noise = input.new_empty(b, 1, h * f, w * f).normal_()
            return self.conv(input, style, bias=self.activate.bias, noise=noise,
                             noise_weight=self.noise.weight, activate=True)
        return self.conv(input, style, bias=self.activate.bias, activate=True)


class ToRGB(nn.Module):
    """:406-425"""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        if _tc(input):
            # the RGB skip path stays f32 NCHW (tiny tensors): slice the padded conv output and convert
            out = self.conv(input, style).float().contiguous() + self.bias
        else:
            out = self.conv(input, style, bias=self.bias.view(-1))
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class Attention(nn.Module):
    """:862-901 — queries from the P tokens, keys/values from the (normalised) Z tokens."""

    def __init__(self, in_dim, param_dim, out_dim, lr_mul=1.0, groups=4, compress=4):
        assert out_dim % (groups * compress) == 0
        super().__init__()
        self.in_dim, self.param_dim, self.out_dim = in_dim, param_dim, out_dim
        self.compress, self.groups = compress, groups
        self.planes = out_dim // compress
        self.group_planes = self.planes // groups
        self.scale = self.planes ** -0.5
        self.q_transform = EqualLinear(param_dim, self.planes, lr_mul=lr_mul)
        self.k_transform = EqualLinear(in_dim, self.planes, lr_mul=lr_mul)
        self.v_transform = EqualLinear(in_dim, self.planes, lr_mul=lr_mul)
        self.proj = EqualLinear(self.planes, out_dim, lr_mul=lr_mul)

    def forward(self, attention, op_param, return_similarity=False):
        n, l, _ = attention.shape
        m = op_param.shape[1]
        q = self.q_transform(op_param)
        k = self.k_transform(attention)
        v = self.v_transform(attention)
        if m == l and self.groups == 4:
            stacked, similarity = op.attn_core(q, k, v)
        else:  # general token counts: literal form of :888-894
            g, gp = self.groups, self.group_planes
            qh = q.reshape(n, m, g, gp).permute(0, 2, 3, 1)
            kh = k.reshape(n, l, g, gp).permute(0, 2, 3, 1)
            vh = v.reshape(n, l, g, gp).permute(0, 2, 3, 1)
            similarity = F.softmax(torch.einsum("abcd,abce->abde", qh, kh) * self.scale, dim=3)
            sv = torch.einsum("abcd,abed->abec", similarity, vh)
            stacked = sv.reshape(n, self.planes, l).permute(0, 2, 1)
        output = self.proj(stacked)
        return (output, similarity) if return_similarity else output


class AttentionBlock(nn.Module):
    """:904-936 — joint (tokens x channels) LayerNorm, no affine."""

    def __init__(self, in_dim, param_dim, out_dim, lr_mul=1.0, groups=4):
        super().__init__()
        self.in_dim, self.out_dim, self.param_dim = in_dim, out_dim, param_dim
        self.atten = Attention(in_dim, param_dim, out_dim, lr_mul=lr_mul, groups=groups)
        self.mlp = nn.Sequential(EqualLinear(out_dim, out_dim, lr_mul=lr_mul), nn.GELU(),
                                 EqualLinear(out_dim, out_dim, lr_mul=lr_mul))
        if out_dim != in_dim:
            self.proj = EqualLinear(in_dim, out_dim, lr_mul=lr_mul)

    def fused_params(self):
        """This block's parameters as te_attn_stack's table entry (include/te_b200.h: te_attn_block)."""
        a, has_proj = self.atten, self.out_dim != self.in_dim
        return {"w_proj": self.proj.weight if has_proj else None, "b_proj": self.proj.bias if has_proj else None,
                "w_q": a.q_transform.weight, "b_q": a.q_transform.bias, "w_k": a.k_transform.weight,
                "b_k": a.k_transform.bias, "w_v": a.v_transform.weight, "b_v": a.v_transform.bias,
                "w_o": a.proj.weight, "b_o": a.proj.bias, "w_m1": self.mlp[0].weight, "b_m1": self.mlp[0].bias,
                "w_m2": self.mlp[2].weight, "b_m2": self.mlp[2].bias, "in_dim": self.in_dim,
                "param_dim": self.param_dim}

    def forward(self, x, op_param, return_similarity=False):
        normed = F.layer_norm(x, x.shape[1:])
        attention, similarity = self.atten(normed, op_param, return_similarity=True)
        x = (self.proj(x) if self.out_dim != self.in_dim else x) + attention
        x = x + self.mlp(F.layer_norm(x, x.shape[1:]))
        return (x, similarity) if return_similarity else x


class Generator(nn.Module):
    """:428-728"""

    def __init__(self, size, style_dim, param_dim, token_dim, channel_multiplier=2,
                 blur_kernel=[1, 3, 3, 1], lr_mlp=0.01, layer_noise_injection=False,
                 use_spatial_mapping=True, num_region=1, n_trans=4, pixel_norm_op_dim=2,
                 no_trans=False):
        super().__init__()
        self.size = size
        self.lr_mlp = lr_mlp
        self.n_trans = n_trans
        self.no_trans = no_trans
        self.style_dim = style_dim
        self.param_dim = param_dim
        self.token_dim = token_dim
        self.layer_noise_injection = layer_noise_injection
        self.style = None
        self.use_spatial_mapping = use_spatial_mapping
        self.num_region = num_region
        self.num_spatial_mapping = int(16 / num_region)
        self.num_style_mapping = self.num_spatial_mapping
        self.pixel_norm_op_dim = pixel_norm_op_dim

        if self.use_spatial_mapping:
            self.spatial_mapping_network = self.spatial_mapping()
        self.style_mapping_network = self.style_mapping()

        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                         128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                         512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        self.adjust_style = EqualLinear(in_dim=16, out_dim=self.token_dim)
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim,
                                blur_kernel=blur_kernel,
                                layer_noise_injection=self.layer_noise_injection)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel,
                                         layer_noise_injection=self.layer_noise_injection))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim,
                                         blur_kernel=blur_kernel,
                                         layer_noise_injection=self.layer_noise_injection))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        self.register_buffer("token", torch.eye(self.token_dim))
        self.register_buffer("token_spatial", torch.eye(16))
        self.trans_interact = not self.no_trans
        if self.trans_interact:
            self.interact = self.interaction_network()

    # -- sub-network builders (names are part of the reference API, :546-577)
    def _mapping(self, count):
        layers = [PixelNorm(self.pixel_norm_op_dim)]
        layers += [EqualLinear(in_dim=self.style_dim, out_dim=self.style_dim, lr_mul=self.lr_mlp,
                               activation="fused_lrelu") for _ in range(count)]
        return nn.Sequential(*layers)

    def style_mapping(self):
        return self._mapping(self.num_style_mapping)

    def spatial_mapping(self):
        return self._mapping(self.num_spatial_mapping)

    def interaction_network(self):
        blocks = [AttentionBlock(in_dim=self.style_dim + 16, param_dim=self.param_dim + 16,
                                 out_dim=self.style_dim, lr_mul=self.lr_mlp)]
        blocks += [AttentionBlock(in_dim=self.style_dim, param_dim=self.param_dim,
                                  out_dim=self.style_dim, lr_mul=self.lr_mlp)
                   for _ in range(1, self.n_trans)]
        return nn.Sequential(*blocks)

    def make_noise(self):
        """:579-588 (the reference hard-codes device='cuda')"""
        device = self.token.device if self.token.is_cuda else "cuda"
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def _prepare_styles(self, latent):
        """bf16 route: the per-layer style work (modulation linear, demodulation coefficients, per-sample
        weights; ~11 small kernels per layer forward, ~3x that backward) depends only on `latent` and the
        weights, so it is issued up front on a SIDE stream and overlaps the convolutions on the main stream —
        in the backward pass too, since autograd runs each node on its forward stream.  Each layer waits on its
        own event.  Disabled with TE_STYLE_STREAM=0."""
        layers = [(self.conv1.conv, 0, 4), (self.to_rgb1.conv, 1, 4)]
        i, res = 1, 4
        for conv1, conv2, to_rgb in zip(self.convs[::2], self.convs[1::2], self.to_rgbs):
            layers.append((conv1.conv, i, res))      # the upsampling conv sees the LOW-resolution input
            res *= 2
            layers.append((conv2.conv, i + 1, res))
            layers.append((to_rgb.conv, i + 2, res))
            i += 2
        for m, _, _ in layers:
            m._prepared = None
        if not latent.is_cuda or os.environ.get("TE_STYLE_STREAM", "1") == "0":
            return
        main = torch.cuda.current_stream(latent.device)
        side = _side_stream(latent.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            # every layer's style modulation (EqualLinear(style_dim, Cin, bias_init=1), :283) in ONE grouped launch
            mods = op.grouped_linear([(latent[:, idx], m.modulation.weight, m.modulation.bias, m.modulation.scale,
                                       m.modulation.lr_mul, False) for m, idx, _ in layers],
                                     tf32=_PRECISION == "bf16")
        for (m, idx, r), s in zip(layers, mods):
            m.prepare(latent[:, idx], r * r, side, main, s=s)

    def _map_columns(self, code, network, count):
        """:626-646 in ONE launch (op.mapping_columns -> te_linear_grouped): PixelNorm, the `count` per-column
        EqualLinears, bias and fused leaky ReLU; no stacked weight copy, no library GEMM.  Returns [B,D,C]."""
        layers = [network[i + 1] for i in range(count)]
        fused_norm = network[0].pixel_norm_op_dim in (1, -2) and code.shape[1] <= 512
        if not fused_norm:
            code = network[0](code)
        return op.mapping_columns(code.float(), [l.weight for l in layers], [l.bias for l in layers], layers[0].scale,
                                  layers[0].lr_mul, tf32=_PRECISION == "bf16", pixel_norm=fused_norm)

    def forward(self, style, op_param, return_latents=False, input_is_latent=False, noise=None,
                randomize_noise=True, return_style=False, return_p_latent=False,
                return_only_style=False, return_only_style_latent=False,
                return_only_mapped_p=False, return_only_mapped_z=False, use_spatial_mapping=True,
                use_style_mapping=True, trans_interact=True, return_mapped_codes=False):
        if self.no_trans:
            trans_interact = False
        if input_is_latent:  # :618-621
            use_spatial_mapping, use_style_mapping, trans_interact = True, False, False

        if use_spatial_mapping:
            spatialcode = self._map_columns(op_param, self.spatial_mapping_network, self.num_spatial_mapping)
        else:
            spatialcode = op_param
        if use_style_mapping:
            stylecode = self._map_columns(style, self.style_mapping_network, self.num_style_mapping)
        else:
            stylecode = style

        if return_mapped_codes:
            return stylecode, spatialcode
        if return_only_mapped_p:
            return spatialcode
        if return_only_mapped_z:
            return stylecode

        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]

        stylecode = stylecode.permute(0, 2, 1)      # [B, 16, 512]
        spatialcode = spatialcode.permute(0, 2, 1)  # [B, 16, 512]
        if trans_interact:
            eye = self.token_spatial.repeat(stylecode.size(0), 1, 1)
            x0, p0 = torch.cat([stylecode, eye], 2), torch.cat([spatialcode, eye], 2)
            blocks = [blk.fused_params() for blk in self.interact]
            if os.environ.get("TE_ATTN_STACK", "1") != "0" and op.attn_stack_supported(x0, p0, spatialcode, blocks):
                # :668-679 in one launch (backward: two) instead of ~35 per block
                x = op.attn_stack(x0, p0, spatialcode if self.n_trans > 1 else None, blocks, self.lr_mlp,
                                  tf32=_PRECISION == "bf16")
            else:
                x = self.interact[0](x0, p0)
                for i in range(1, self.n_trans):
                    x = self.interact[i](x, spatialcode)

        if self.no_trans:
            latent = self.adjust_style(stylecode.permute(0, 2, 1)).permute(0, 2, 1)
        elif not input_is_latent:
            if not trans_interact:
                raise NameError("trans_interact=False without no_trans/input_is_latent leaves the "
                                "latent undefined (the reference raises here too, :686)")
            latent = self.adjust_style(x.permute(0, 2, 1)).permute(0, 2, 1)  # [B, T, 512]
        else:
            latent = style

        if return_only_style_latent or return_only_style:
            return latent

        batch = spatialcode.shape[0]
        out = spatialcode.permute(0, 2, 1).reshape(batch, 512, 4, 4)
        if _tc_mode():
            out = _to_cl(out)
            self._prepare_styles(latent)
        out = self.conv1(out, latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1])
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2],
                                                        noise[1::2], noise[2::2], self.to_rgbs):
            out = conv1(out, latent[:, i], noise=noise1)
            out = conv2(out, latent[:, i + 1], noise=noise2)
            skip = to_rgb(out, latent[:, i + 2], skip)
            i += 2
        image = skip

        if return_style:
            return image, latent
        if return_p_latent:
            return image, spatialcode
        if return_latents:
            return image, latent, None
        return image, None, None


class ConvLayer(nn.Sequential):
    """:731-777"""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False,
                 blur_kernel=[1, 3, 3, 1], bias=True, activate=True):
        layers = []
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            layers.append(Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride = 2
            self.padding = 0
        else:
            stride = 1
            self.padding = kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding,
                                  stride=stride, bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)

    def forward(self, input):
        if not _tc(input):
            return super().forward(input)
        return self.forward_tc(input)

    def forward_tc_carry(self, input):
        """(layer(input), input_alias) for a layer that is exactly EqualConv2d + FusedLeakyReLU: the alias feeds
        the input's second consumer so that the two gradient contributions are summed inside this layer's
        data-gradient kernel (tc.TcConvBiasActCarry).  Falls back to (forward_tc(input), input)."""
        mods = list(self)
        if (len(mods) == 2 and isinstance(mods[0], EqualConv2d) and isinstance(mods[1], FusedLeakyReLU)
                and mods[1].bias is not None and mods[0].bias is None and mods[0].weight.shape[0] % 8 == 0
                and mods[1].negative_slope == 0.2 and mods[0].stride == 1 and torch.is_grad_enabled()
                and input.requires_grad):
            m, act = mods
            return tc.conv2d_bias_act_carry(input, m._tc_weight(input), act.bias, stride=1, wscale=m.scale,
                                            gain=act.scale)
        return self.forward_tc(input), input

    def forward_tc(self, input, out_mul=1.0, residual=None):
        """bf16 tensor-core route, returns layer(input) * out_mul + residual.  [Blur] -> EqualConv2d +
        FusedLeakyReLU run as ONE kernel (bias and activation in the convolution epilogue); the equalised-lr
        scale, `out_mul` (folded into the activation gain or the weight scale) and the residual sum cost no
        pass of their own."""
        mods = list(self)
        x = input
        i = 0
        stride1 = False
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(m, Blur) and isinstance(nxt, EqualConv2d) and nxt.weight.shape[2] == 1
                    and nxt.stride == 2 and nxt.padding == 0):
                # blur + stride-2 1x1 conv (the ResBlock skip): evaluate the FIR only where the conv reads it
                # (upfirdn2d down=2, a quarter of the outputs) and run the 1x1 conv at stride 1; the backward
                # becomes an up=2 FIR of the compact gradient instead of a zero-filled full-resolution one
                x = upfirdn2d(x, m.kernel, down=2, pad=m.pad)
                stride1 = True
                i += 1
                continue
            last = i + (2 if isinstance(nxt, FusedLeakyReLU) else 1) >= len(mods)
            if (isinstance(m, EqualConv2d) and isinstance(nxt, FusedLeakyReLU) and nxt.bias is not None
                    and m.bias is None and m.weight.shape[0] % 8 == 0 and nxt.negative_slope == 0.2
                    and (residual is None or not last)):
                gain = nxt.scale * (out_mul if last else 1.0)
                x = tc.conv2d_bias_act(x, m._tc_weight(x), nxt.bias, stride=m.stride, wscale=m.scale, gain=gain)
                if last:
                    out_mul = 1.0
                i += 2
            elif isinstance(m, EqualConv2d) and m.bias is None and nxt is None and m.weight.shape[0] % 8 == 0:
                wscale = m.scale * out_mul
                stride = 1 if stride1 else m.stride
                if residual is not None:
                    x = tc.conv2d_residual(x, m._tc_weight(x), residual, stride=stride, wscale=wscale)
                    residual = None
                else:
                    x = tc.conv2d(x, m._tc_weight(x), stride=stride, wscale=wscale)
                out_mul = 1.0
                i += 1
            else:
                x = m(x)
                i += 1
        if out_mul != 1.0:
            x = x * out_mul
        if residual is not None:
            x = x + residual
        return x


class ResBlock(nn.Module):
    """:780-798"""

    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3, blur_kernel=blur_kernel)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=True, blur_kernel=blur_kernel)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=True, blur_kernel=blur_kernel,
                              bias=False, activate=False)

    def forward(self, input):
        if _tc(input):
            # (conv2(conv1(x)) + skip(x)) / sqrt(2) with the 1/sqrt(2) folded into conv2's activation gain and the
            # skip convolution's weight scale, and the sum taken in the skip convolution's epilogue
            # and the two gradient contributions of `input` summed in conv1's data-gradient epilogue (carry)
            h, input_alias = self.conv1.forward_tc_carry(input)
            out = self.conv2.forward_tc(h, out_mul=1 / _SQRT2)
            return self.skip.forward_tc(input_alias, out_mul=1 / _SQRT2, residual=out)
        out = self.conv2(self.conv1(input))
        return (out + self.skip(input)) / _SQRT2


class Discriminator(nn.Module):
    """:801-859"""

    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                    128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                    512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        convs = [ConvLayer(3, channels[size], 1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 4
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(
            EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
            EqualLinear(channels[4], 1))

    def forward(self, input):
        return self.forward_stacked(input, 1)

    def forward_stacked(self, input, sub_batches):
        """Not in the reference: treats `input` as `sub_batches` independent batches stacked along dim 0 —
        e.g. cat([fake, real]) — so ONE pass gives exactly the logits of separate calls: every layer is
        per-sample except the minibatch standard deviation, which is taken per sub-batch."""
        bf16 = _tc_mode() and input.dtype in (torch.float32, torch.bfloat16)
        # a LEAF image that requires grad is the R1 penalty's input (train_spatial_query.py:212-216): its gradient
        # will be differentiated again, which the twice-differentiable tensor-core ops do without a recomputation
        second_order = input.requires_grad and input.is_leaf and torch.is_grad_enabled()
        if bf16 and input.requires_grad and torch.is_grad_enabled():
            input = _DenseGrad.apply(input)
        # tensor-core modes: the whole conv stack incl. final_conv on tensor cores; statistics and linears in f32
        # RGB is zero-padded to 64 channels: a TMA box whose rows are mostly out of bounds (8 of 64 channels)
        # takes the unit's slow path (measured 4 us per tile); a dense 128-byte row costs 134 MB of extra
        # input but runs at full speed
        first = self.convs[0]
        if (bf16 and input.dtype == torch.float32 and input.shape[1] == 3 and len(first) == 2
                and os.environ.get("TE_FROM_RGB", "1") != "0"
                and not second_order
                and isinstance(first[0], EqualConv2d) and isinstance(first[1], FusedLeakyReLU)
                and first[0].weight.shape[2] == 1 and first[0].bias is None and first[1].bias is not None
                and first[1].negative_slope == 0.2 and first[0].weight.shape[0] in (8, 16, 32, 64, 128, 256)):
            # from-RGB (ConvLayer(3, C, 1), :806-808): K = 3 is a stream, not a GEMM — one dedicated kernel each way
            # straight from the f32 NCHW image (op.FromRGB) instead of a padded tensor-core launch
            out = op.from_rgb(input, first[0].weight, first[1].bias, first[0].scale, first[1].scale, _act_dtype())
            for layer in list(self.convs)[1:]:
                out = layer(out)
        else:
            out = self.convs(_to_cl(input, pad_to=64) if bf16 else input)
        batch, channel, height, width = out.shape
        if batch % sub_batches:
            raise ValueError("batch %d is not divisible into %d sub-batches" % (batch, sub_batches))
        sub = batch // sub_batches
        group = min(sub, self.stddev_group)
        # minibatch standard deviation, one scalar per group column of each sub-batch (:844-852)
        stat_in = out.float().contiguous() if bf16 else out
        grouped = stat_in.view(sub_batches, group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(grouped.var(1, unbiased=False) + 1e-8)          # [S, M, feat, C/feat, H, W]
        stddev = stddev.mean([3, 4, 5], keepdim=True).squeeze(3)            # [S, M, feat, 1, 1]
        stddev = stddev.unsqueeze(1).expand(sub_batches, group, -1, self.stddev_feat, height, width)
        stddev = stddev.reshape(batch, self.stddev_feat, height, width)
        if bf16:
            out = _to_cl(torch.cat([out, stddev.to(out.dtype)], 1))  # 513 -> 520 channels
            out = self.final_conv(out).float().contiguous()
        else:
            out = self.final_conv(torch.cat([out, stddev], 1))
        return self.final_linear(out.view(batch, -1))
