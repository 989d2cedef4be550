"""Tensor-core (tcgen05) convolution engine, bf16 channels-last: forward, data gradient and weight
gradient of every convolution on the path, as twice-differentiable autograd Functions.

Three geometries ("modes"), written for one axis (K taps, p = K // 2):
    s1    y[a]      = sum_k x[a + k - p] . Wsel[k]       stride-1 "same" convolution
    down  y[a]      = sum_k x[2a + k]    . Wsel[k]       stride-2 convolution, no padding
    down1 y[a]      = sum_k x[2a + k - 1]. Wsel[k]       stride-2 convolution, padding 1 (inference only: the pSp heads)
    up    y[2i + k] += x[i]              . Wsel[k]       transposed stride-2 convolution (full)
where Wsel[k] = W[k] or W[K-1-k] (`flip`) taken as [Cout, Cin] or transposed (`transposed`) slices of
the master weight W [O, I, K, K].  The set is closed under differentiation:
    adjoint(s1)   = s1   with transposed^, flip^
    adjoint(down) = up   with transposed^
    adjoint(up)   = down with transposed^
and the weight gradient of each mode is one te_conv_wgrad_tc call per launch of the forward geometry.
`up` is issued as one launch per output parity class (polyphase), so no zero-inserted tensor exists.

Two operand precisions, selected by the dtype of the activations:
    bf16 activations   plain bf16 x bf16 -> f32 products (the speed mode)
    f32 activations    SPLIT-OPERAND mode: activations, gradients and weights are stored as `split_planes()` bf16
                       planes hi = bf16(v), mid = bf16(v - hi)(, lo) and the kernels sum the plane-pair products
                       hi*hi + hi*mid + mid*hi (+ mid*mid + hi*lo + lo*hi) in their f32 accumulators: the reference's
                       fp32 F.conv2d arithmetic to ~2^-16 per product with two planes (default; 2^-24 with three)
                       ON TENSOR CORES.  Outputs are f32 channels-last.  `set_split_planes(2|3)` picks the count.

Reference call sites replaced: F.conv2d / F.conv_transpose2d in model_spatial_query.py:177-183,318,327,333
and their autograd gradients.
"""
import os

import torch
from torch.autograd import Function

from . import lib


def _pad8(c):
    return (c + 7) // 8 * 8


_SPLIT_PLANES = 2


def set_split_planes(n):
    """Planes used for f32 activations: 2 (three products, ~2^-16 per product: the default) or 3 (six products,
    twice the cost).  Measured image max-abs error against the reference goldens / the CPU restatement on B200
    (profiles/r02_parity_modes_v2.txt): 2 planes 2.2e-4 at 256^2 and 4.5e-4 at 1024^2; 3 planes 9.6e-5 and 1.9e-4;
    the exact-f32 SIMT engine 4.2e-5 and 9.7e-5 — all inside the 1e-3 bar."""
    global _SPLIT_PLANES
    if n not in (2, 3):
        raise ValueError("split planes must be 2 or 3")
    _SPLIT_PLANES = n


def get_split_planes():
    return _SPLIT_PLANES


def nseg_for(x):
    """Operand planes the kernels use for activations of x's dtype."""
    return _SPLIT_PLANES if x.dtype == torch.float32 else 1


def split_planes(x, nseg, scale=None):
    """f32 [B, C, H, W] (any layout) -> bf16 planes [nseg, B, H, W, C] of x (* scale[b, c] when given)."""
    lib.require_cuda(x, scale)
    x = x.contiguous(memory_format=torch.channels_last)
    b, c, h, w = x.shape
    if c % 8:
        raise RuntimeError("tensor-core conv needs channel counts that are multiples of 8 (got %d)" % c)
    out = torch.empty((nseg, b, h, w, c), dtype=torch.bfloat16, device=x.device)
    sf = None if scale is None else scale.to(torch.float32).contiguous()
    lib.split_bf16(out, x, sf, b, h * w, c, nseg)
    return out


def _split_torch(v, nseg):
    """f32 tensor -> stacked bf16 planes [nseg, ...] with torch ops (small per-sample weight tensors)."""
    planes, r = [], v.contiguous()
    for i in range(nseg):
        h = r.to(torch.bfloat16)
        planes.append(h)
        if i + 1 < nseg:
            r = r - h.float()
    return torch.stack(planes).contiguous()


class Mode:
    """Geometry of one launch family."""
    __slots__ = ("kind", "k", "transposed", "flip", "out_hw")

    def __init__(self, kind, k, transposed=False, flip=False, out_hw=None):
        self.kind, self.k, self.transposed, self.flip = kind, k, transposed, flip
        self.out_hw = None if out_hw is None else (int(out_hw[0]), int(out_hw[1]))

    def adjoint(self, in_hw):
        if self.kind == "s1":
            return Mode("s1", self.k, not self.transposed, not self.flip, in_hw)
        if self.kind == "down":
            return Mode("up", self.k, not self.transposed, self.flip, in_hw)
        if self.kind == "down1":
            raise NotImplementedError("the padded stride-2 geometry is forward-only")
        return Mode("down", self.k, not self.transposed, self.flip, in_hw)

    def output_hw(self, hin, win):
        k = self.k
        if self.out_hw is not None:
            return self.out_hw
        if self.kind == "s1":
            return hin, win
        if self.kind == "down":
            return (hin - k) // 2 + 1, (win - k) // 2 + 1
        if self.kind == "down1":
            return (hin + 2 - k) // 2 + 1, (win + 2 - k) // 2 + 1
        return (hin - 1) * 2 + k, (win - 1) * 2 + k

    def launches(self, hin, win, hout, wout):
        """List of (taps, in_stride, out_stride, off_y, off_x, grid_h, grid_w); taps = [(dy, dx, widx)]."""
        k = self.k
        sel = (lambda t: k - 1 - t) if self.flip else (lambda t: t)
        widx = lambda ky, kx: sel(ky) * k + sel(kx)  # noqa: E731
        if self.kind == "s1":
            taps = [(ky - k // 2, kx - k // 2, widx(ky, kx)) for ky in range(k) for kx in range(k)]
            return [(taps, 1, 1, 0, 0, hout, wout)]
        if self.kind == "down":
            taps = [(ky, kx, widx(ky, kx)) for ky in range(k) for kx in range(k)]
            return [(taps, 2, 1, 0, 0, hout, wout)]
        if self.kind == "down1":
            taps = [(ky - 1, kx - 1, widx(ky, kx)) for ky in range(k) for kx in range(k)]
            return [(taps, 2, 1, 0, 0, hout, wout)]
        out = []
        for ry in (0, 1):
            for rx in (0, 1):
                gh, gw = (hout - ry + 1) // 2, (wout - rx + 1) // 2
                taps = [(-(ky - ry) // 2, -(kx - rx) // 2, widx(ky, kx))
                        for ky in range(ry, k, 2) for kx in range(rx, k, 2)]
                if taps and gh > 0 and gw > 0:
                    out.append((taps, 1, 2, ry, rx, gh, gw))
        return out

    def covers_output(self):
        """False when some output positions receive no tap (k=1 transposed: only even positions)."""
        return not (self.kind == "up" and self.k == 1)


def _scaled_copy(dst, src, scale):
    """dst (bf16 view) = src (f32, any strides) * scale: permute + scale + f32->bf16 in ONE kernel."""
    if scale == 1.0:
        dst.copy_(src)
    else:
        torch.mul(src, scale, out=dst)


class PackCache:
    """Persistent bf16 tap-major copies of SHARED convolution weights.

    Without it every convolution call repacks its f32 master weight (permute + scale + down-cast: 111 launches, ~1.1 ms
    per training iteration).  A trainer registers its parameters, refreshes the copies once after each optimiser
    step — ONE table-driven launch (te_pack_weights_tc) for all weights of a model, both orientations — and
    `pack_weight` hands out the stored copy.  Only whole registered parameters hit (matched by storage address and
    shape); padded / sliced / per-sample / second-order weights are packed per call as before.  Whoever changes a
    registered weight outside the trainer's optimiser (load_state_dict, manual edits) must call `refresh()`."""

    def __init__(self):
        self._registered = {}   # data_ptr -> (O, I, K, K)
        self._entries = {}      # (data_ptr, scale, nseg) -> [src, dst_n, dst_t, version of src when last packed]
        self._keepalive = []    # storages of the registered parameters: their addresses cannot be recycled while
                                # this cache lives, so a data_ptr match is a match of the parameter itself

    def register(self, params):
        for p in params:
            shape = tuple(p.shape[1:]) if (p.dim() == 5 and p.shape[0] == 1) else tuple(p.shape)
            if len(shape) == 4 and shape[2] == shape[3] and shape[2] in (1, 3) and p.dtype == torch.float32:
                self._registered[p.data_ptr()] = shape
                self._keepalive.append(p.untyped_storage())

    def lookup(self, w, transposed, scale, nseg=1):
        if w.dim() != 4 or self._registered.get(w.data_ptr()) != tuple(w.shape) or not w.is_contiguous():
            return None
        key = (w.data_ptr(), float(scale), int(nseg))
        e = self._entries.get(key)
        if e is None:
            e = self._entries[key] = [w.detach(), None, None, w._version]
        slot = 2 if transposed else 1
        stale = e[3] != w._version
        if e[slot] is None:
            o, i, k, _ = w.shape
            po, pi = _pad8(o), _pad8(i)
            shape = (k * k, pi, po) if transposed else (k * k, po, pi)
            e[slot] = torch.zeros(shape if nseg == 1 else (nseg,) + shape, dtype=torch.bfloat16, device=w.device)
            stale = True
        if stale:
            # new slot, or the weight changed through torch (load_state_dict, copy_, a torch optimiser: those bump the
            # version counter; the trainer's own fused Adam writes through raw pointers and calls refresh() instead):
            # re-pack EVERY existing orientation, so that none is left behind a matching version number
            lib.pack_weights_tc([(e[0], e[1], e[2], float(scale), int(nseg))])
        e[3] = w._version
        return e[slot]

    def refresh(self, lo=None, hi=None):
        """Re-pack every stored copy (of the weights whose storage lies in [lo, hi) when given): one launch per 64."""
        tasks = [(e[0], e[1], e[2], key[1], key[2]) for key, e in self._entries.items()
                 if (lo is None or lo <= key[0] < hi) and (e[1] is not None or e[2] is not None)]
        lib.pack_weights_tc(tasks)


_PACK_CACHE = None


def set_pack_cache(cache):
    """Install (or with None remove) the PackCache consulted by `pack_weight`."""
    global _PACK_CACHE
    _PACK_CACHE = cache


class use_pack_cache:
    """Context manager: `pack_weight` consults `cache` inside the block and whatever was installed before after it.
    A Trainer wraps its own phases in this, so two trainers (or a trainer and free-standing modules) in one process
    never see each other's cached copies."""

    def __init__(self, cache):
        self.cache = cache

    def __enter__(self):
        global _PACK_CACHE
        self.prev, _PACK_CACHE = _PACK_CACHE, self.cache
        return self.cache

    def __exit__(self, *exc):
        global _PACK_CACHE
        _PACK_CACHE = self.prev
        return False


def pack_weight(w, transposed, scale=1.0, nseg=1):
    """Master weight [O, I, K, K] (f32) times `scale` -> bf16 slices [K*K, Cout, Cin] (Cin contiguous), channel
    counts padded to multiples of 8 with zeros.  A leading batch dimension ([B, O, I, K, K] ->
    [B, K*K, Cout, Cin]) gives per-sample weights.  nseg = 2 or 3: split-operand planes, stacked in front
    ([nseg, K*K, Cout, Cin] / [nseg, B, K*K, Cout, Cin])."""
    lib.require_cuda(w)
    if _PACK_CACHE is not None and w.dim() == 4:
        hit = _PACK_CACHE.lookup(w, transposed, scale, nseg)
        if hit is not None:
            return hit
    if w.dim() == 5:
        b, o, i, k, _ = w.shape
        if o % 8 or i % 8:
            raise RuntimeError("per-sample tensor-core weights need channel counts that are multiples of 8")
        if transposed:
            w = w.transpose(1, 2)
            o, i = i, o
        if nseg > 1:
            v = (w.detach().to(torch.float32) * scale).permute(0, 3, 4, 1, 2).reshape(b, k * k, o, i)
            return _split_torch(v, nseg)
        out = torch.empty((b, k * k, o, i), dtype=torch.bfloat16, device=w.device)
        _scaled_copy(out.view(b, k, k, o, i), w.permute(0, 3, 4, 1, 2), scale)
        return out
    if nseg > 1:  # one table-driven launch writes the planes of either orientation
        o, i, k, _ = w.shape
        po, pi = _pad8(o), _pad8(i)
        out = torch.zeros((nseg, k * k, pi, po) if transposed else (nseg, k * k, po, pi), dtype=torch.bfloat16,
                          device=w.device)
        src = w.detach().to(torch.float32).contiguous()
        lib.pack_weights_tc([(src, None if transposed else out, out if transposed else None, float(scale), nseg)])
        return out
    o, i, k, _ = w.shape
    if transposed:
        w = w.transpose(0, 1)
        o, i = i, o
    po, pi = _pad8(o), _pad8(i)
    alloc = torch.empty if (po == o and pi == i) else torch.zeros
    out = alloc((k * k, po, pi), dtype=torch.bfloat16, device=w.device)
    _scaled_copy(out.view(k, k, po, pi)[:, :, :o, :i], w.permute(2, 3, 0, 1), scale)
    return out


def _desc(x, cout, hout, wout, launch, w_slices, act=0, out_f32=0, per_sample=False, act_gain=0.0,
          wgrad_alpha=0.0, residual=None, slope=None, split=1):
    taps, ist, ost, oy, ox, gh, gw = launch
    b, cin, hin, win = x.shape
    d = lib.TcConvDesc()
    d.batch, d.hin, d.win, d.cin = b, hin, win, cin
    d.hout, d.wout, d.cout = hout, wout, cout
    d.ntaps = len(taps)
    for t, (dy, dx, wi) in enumerate(taps):
        d.tap_dy[t], d.tap_dx[t], d.tap_w[t] = dy, dx, wi
    d.w_slices = w_slices
    d.in_stride, d.out_stride, d.out_off_y, d.out_off_x = ist, ost, oy, ox
    d.grid_h, d.grid_w = gh, gw
    d.act, d.out_f32 = act, out_f32
    d.w_bstride = w_slices * cout * cin if per_sample else 0
    d.act_gain, d.wgrad_alpha = act_gain, wgrad_alpha
    d.residual = residual.data_ptr() if residual is not None else None
    d.slope = slope.data_ptr() if slope is not None else None
    d.split = split
    d.py_refs = (residual, slope)  # keeps the two tensors alive for the launch (and lets tests/emu.py read them)
    return d


def _cl_act(x):
    """Activations for the kernels: (operand tensor, nseg).  bf16 -> the channels-last tensor itself, 1;
    f32 -> its split-operand planes [nseg, B, H, W, C], nseg."""
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("tensor-core conv needs bf16 or f32 activations, got %s" % x.dtype)
    if x.shape[1] % 8:
        raise RuntimeError("tensor-core conv needs channel counts that are multiples of 8 (got %d)" % x.shape[1])
    if x.dtype == torch.float32:
        return split_planes(x, _SPLIT_PLANES), _SPLIT_PLANES
    return x.contiguous(memory_format=torch.channels_last), 1


def conv_raw(x, wp, mode, out_scale=None, bias=None, act=False, act_gain=0.0, residual=None, slope=None,
             in_scale=None):
    """x [B, Cin, H, W] bf16 (or f32: split-operand mode) channels-last, wp = pack_weight(..., nseg=nseg_for(x)).
    Returns a channels-last tensor of x's dtype: act(conv * out_scale + bias) * act_gain + residual.
    act: False/0, True/1 = leaky 0.2 (gain sqrt 2 by default), 2 = leaky 0.01 (gain 1 by default), 3 = PReLU with the
    f32 per-channel `slope` [Cout].  in_scale [B, Cin] (f32 activations only): x is modulated while it is split."""
    lib.require_cuda(x, wp, out_scale, bias, residual, slope, in_scale)
    if (int(act) == 3) != (slope is not None):
        raise RuntimeError("conv_tc: act=3 (PReLU) and `slope` go together")
    if slope is not None:
        slope = slope.to(torch.float32).contiguous()
    b, cin, hin, win = x.shape
    dtype = x.dtype
    if in_scale is not None:
        if dtype != torch.float32:
            raise RuntimeError("conv_tc: in_scale is fused into the f32 operand split only")
        xo, nseg = split_planes(x, _SPLIT_PLANES, in_scale), _SPLIT_PLANES
    else:
        xo, nseg = _cl_act(x)
    per_sample = wp.dim() == (4 if nseg == 1 else 5)
    if wp.dim() not in ((3, 4) if nseg == 1 else (4, 5)) or (nseg > 1 and wp.shape[0] != nseg):
        raise RuntimeError("conv_tc: weight %s was not packed for %d operand plane(s)" % (tuple(wp.shape), nseg))
    cout = wp.shape[-2]
    if wp.shape[-1] != cin or (per_sample and wp.shape[-4] != b):
        raise RuntimeError("conv_tc: weight %s does not match activations %s" % (tuple(wp.shape), tuple(x.shape)))
    hout, wout = mode.output_hw(hin, win)
    y = torch.empty((b, cout, hout, wout), dtype=dtype, device=x.device, memory_format=torch.channels_last)
    if residual is not None:
        if residual.shape != y.shape or residual.dtype != dtype or not mode.covers_output():
            raise RuntimeError("conv_tc: residual must be %s of the output's shape %s" % (dtype, tuple(y.shape)))
        residual = residual.contiguous(memory_format=torch.channels_last)
    if not mode.covers_output():
        y.zero_()
    f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()  # noqa: E731
    osc, bi = f32(out_scale), f32(bias)
    for launch in mode.launches(hin, win, hout, wout):
        lib.conv_tc(y, xo, wp, osc, bi, _desc(x, cout, hout, wout, launch, wp.shape[-3], int(act),
                                              out_f32=int(dtype == torch.float32), per_sample=per_sample,
                                              act_gain=act_gain, residual=residual, slope=slope, split=nseg))
    return y


_WGRAD_WS = {}   # (shape, device, stream) -> all-zero f32 accumulator handed back zeroed by te_wgrad_unpack(clear=1)


def _wgrad_workspace(shape, device):
    """The tap-major f32 accumulator of one weight-gradient launch.  CUDA: a persistent buffer per (shape, stream)
    that is all-zero between uses — te_wgrad_unpack reads it and writes the zeros back in the same pass, so neither a
    memset nor an allocation precedes the ~120 weight-gradient launches of an iteration.  (Kernels on one stream run in
    order, so consecutive launches may share it; TE_WGRAD_WS_CACHE=0 allocates a fresh zeroed buffer per call.)"""
    if device.type != "cuda" or os.environ.get("TE_WGRAD_WS_CACHE", "1") == "0":
        return torch.zeros(shape, dtype=torch.float32, device=device), False
    key = (shape, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WGRAD_WS.get(key)
    if ws is None:
        ws = _WGRAD_WS[key] = torch.zeros(shape, dtype=torch.float32, device=device)
    return ws, True


def clear_wgrad_workspaces():
    """Drop the cached weight-gradient accumulators (a new Trainer / a new set of CUDA graphs starts from fresh ones
    rather than from buffers that live in another graph's memory pool)."""
    _WGRAD_WS.clear()


def wgrad_raw(g, x, mode, w_shape, scale=1.0):
    """Gradient w.r.t. the master weight [O, I, K, K] (f32) of y = conv(x, W * scale; mode) given g = dL/dy."""
    lib.require_cuda(g, x)
    if g.dtype != x.dtype:
        g = g.to(x.dtype)
    b, cin, hin, win = x.shape
    cout, hout, wout = g.shape[1], g.shape[2], g.shape[3]
    (go, nseg), (xo, _) = _cl_act(g), _cl_act(x)
    k = mode.k
    per_sample = len(w_shape) == 5
    shape = (b, k * k, cout, cin) if per_sample else (k * k, cout, cin)
    gw, cached = _wgrad_workspace(shape, x.device)
    o, i = w_shape[-4], w_shape[-3]
    out = torch.empty((b, o, i, k, k) if per_sample else (o, i, k, k), dtype=torch.float32, device=x.device)
    try:
        for launch in mode.launches(hin, win, hout, wout):
            lib.conv_wgrad_tc(gw, go, xo, _desc(x, cout, hout, wout, launch, k * k, per_sample=per_sample,
                                                wgrad_alpha=scale, split=nseg))
        # tap-major accumulator -> [O, I, k, k] (taps innermost) in one coalesced pass that also re-zeroes it
        lib.wgrad_unpack(out, gw, b if per_sample else 1, o, i, k * k, cout, cin, mode.transposed, cached)
    except Exception:
        if cached:   # a half-finished accumulation must not be served again
            _WGRAD_WS.clear()
        raise
    if cached and (o < (cin if mode.transposed else cout) or i < (cout if mode.transposed else cin)):
        gw.zero_()   # padded rows / columns are not visited by the unpack (ToRGB's 3 -> 8 channels): clear them too
    return out


class TcConv(Function):
    """y = conv(x, W * wscale; mode) on tensor cores; W is the f32 master weight [O, I, K, K] and `wscale` a
    python float (the equalised-lr factor) applied while the weight is repacked to bf16."""

    @staticmethod
    def forward(ctx, x, w, mode, wscale=1.0):
        ctx.save_for_backward(x, w)
        ctx.mode, ctx.wscale = mode, wscale
        return conv_raw(x, pack_weight(w, mode.transposed, wscale, nseg_for(x)), mode)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = TcConv.apply(gy, w, mode.adjoint((x.shape[2], x.shape[3])), wscale)
            if gx.shape[1] != x.shape[1]:
                gx = gx[:, :x.shape[1]]
        if ctx.needs_input_grad[1]:
            gw = TcWeightGrad.apply(x, gy, mode, tuple(w.shape), wscale)
        return gx, gw, None, None


class TcWeightGrad(Function):
    """gW = wscale * wgrad(x, gy; mode): bilinear, so its backward is two TcConv calls."""

    @staticmethod
    def forward(ctx, x, gy, mode, w_shape, wscale=1.0):
        ctx.save_for_backward(x, gy)
        ctx.mode, ctx.wscale = mode, wscale
        return wgrad_raw(gy, x, mode, w_shape, wscale)

    @staticmethod
    def backward(ctx, ggw):
        x, gy = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        g_x = g_gy = None
        if ctx.needs_input_grad[0]:
            g_x = TcConv.apply(gy, ggw, mode.adjoint((x.shape[2], x.shape[3])), wscale)
        if ctx.needs_input_grad[1]:
            g_gy = TcConv.apply(x, ggw, Mode(mode.kind, mode.k, mode.transposed, mode.flip,
                                             (gy.shape[2], gy.shape[3])), wscale)
        return g_x, g_gy, None, None, None


class TcConvBiasAct(Function):
    """out = leaky_relu(conv(x, W * wscale; mode) + bias, 0.2) * gain with bias and activation applied in the
    convolution kernel's epilogue (no separate bias-act pass over the activation).  Backward is the
    composition of the differentiable pieces: the masked gradient comes from FusedLeakyReLUFunctionBackward
    (sign taken from the saved OUTPUT, like the reference op, utils/op/fused_act.py:27-29), then the usual
    data / weight gradient kernels — so second order works exactly as for the unfused ops."""

    @staticmethod
    def forward(ctx, x, w, bias, mode, wscale=1.0, gain=2 ** 0.5):
        if not gain > 0:
            raise ValueError("TcConvBiasAct: the gain must be positive (the backward mask is the output's sign)")
        out = conv_raw(x, pack_weight(w, mode.transposed, wscale, nseg_for(x)), mode, bias=bias, act=True, act_gain=gain)
        ctx.save_for_backward(x, w, out)
        ctx.mode, ctx.wscale, ctx.gain = mode, wscale, gain
        ctx.bias_dtype = bias.dtype
        return out

    @staticmethod
    def backward(ctx, g_out):
        from .op import FusedLeakyReLUFunctionBackward
        x, w, out = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        g_y, g_b = FusedLeakyReLUFunctionBackward.apply(g_out, out, True, 0.2, ctx.gain, ctx.bias_dtype)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = TcConv.apply(g_y, w, mode.adjoint((x.shape[2], x.shape[3])), wscale)
            if gx.shape[1] != x.shape[1]:
                gx = gx[:, :x.shape[1]]
        if ctx.needs_input_grad[1]:
            gw = TcWeightGrad.apply(x, g_y, mode, tuple(w.shape), wscale)
        return gx, gw, (g_b if ctx.needs_input_grad[2] else None), None, None, None


class TcConvBiasActCarry(Function):
    """TcConvBiasAct that also hands its INPUT on: returns (out, x_alias).  A second consumer of x (the ResBlock
    skip branch, model_spatial_query.py:795-797) reads x_alias instead of x, so this node receives BOTH gradient
    contributions of x and forms their sum in the data-gradient kernel's epilogue (dgrad + residual) — instead
    of autograd adding two activation-sized tensors in a pass of its own."""

    @staticmethod
    def forward(ctx, x, w, bias, mode, wscale=1.0, gain=2 ** 0.5):
        if not gain > 0:
            raise ValueError("TcConvBiasActCarry: the gain must be positive")
        out = conv_raw(x, pack_weight(w, mode.transposed, wscale, nseg_for(x)), mode, bias=bias, act=True, act_gain=gain)
        ctx.save_for_backward(x, w, out)
        ctx.mode, ctx.wscale, ctx.gain = mode, wscale, gain
        ctx.bias_dtype = bias.dtype
        return out, x.view_as(x)

    @staticmethod
    def backward(ctx, g_out, g_alias):
        from .op import FusedLeakyReLUFunctionBackward
        x, w, out = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        gx = gw = g_b = None
        if g_out is None:
            return g_alias, None, None, None, None, None
        g_y, g_b = FusedLeakyReLUFunctionBackward.apply(g_out, out, True, 0.2, ctx.gain, ctx.bias_dtype)
        if ctx.needs_input_grad[0]:
            adj = mode.adjoint((x.shape[2], x.shape[3]))
            fusable = (g_alias is not None and adj.covers_output() and w.shape[-3] % 8 == 0
                       and g_alias.dtype == g_y.dtype)
            if fusable:
                gx = TcConvResidual.apply(g_y, w, g_alias, adj, wscale)
            else:
                gx = TcConv.apply(g_y, w, adj, wscale)
                if gx.shape[1] != x.shape[1]:
                    gx = gx[:, :x.shape[1]]
                if g_alias is not None:
                    gx = gx + g_alias
        if ctx.needs_input_grad[1]:
            gw = TcWeightGrad.apply(x, g_y, mode, tuple(w.shape), wscale)
        return gx, gw, (g_b if ctx.needs_input_grad[2] else None), None, None, None


class TcConvScaled(Function):
    """y = d[b, cout] * conv(x, W * wscale; mode): the demodulation coefficient of ModulatedConv2d
    (model_spatial_query.py:301-304) applied in the convolution kernel's epilogue instead of a pass of its own over
    the activation.  Backward is a composition of differentiable pieces — g_c = g_y * d (ScaleBC), g_d = sum_pixels
    (g_y * y) / d (DotBC on the saved OUTPUT: y = d * c, and d = rsqrt(.) > 0), then the usual data / weight gradient
    kernels — so second order works like for the unfused ops."""

    @staticmethod
    def forward(ctx, x, w, d, mode, wscale=1.0):
        y = conv_raw(x, pack_weight(w, mode.transposed, wscale, nseg_for(x)), mode, out_scale=d)
        ctx.save_for_backward(x, w, d, y)
        ctx.mode, ctx.wscale = mode, wscale
        return y

    @staticmethod
    def backward(ctx, gy):
        from .op import DotBC, ScaleBC
        x, w, d, y = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        gx = gw = gd = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g_c = ScaleBC.apply(gy, d)
            if ctx.needs_input_grad[0]:
                gx = TcConv.apply(g_c, w, mode.adjoint((x.shape[2], x.shape[3])), wscale)
                if gx.shape[1] != x.shape[1]:
                    gx = gx[:, :x.shape[1]]
            if ctx.needs_input_grad[1]:
                gw = TcWeightGrad.apply(x, g_c, mode, tuple(w.shape), wscale)
        if ctx.needs_input_grad[2]:
            gd = (DotBC.apply(gy, y) / d.to(torch.float32)).to(d.dtype)
        return gx, gw, gd, None, None


class TcConvResidual(Function):
    """y = conv(x, W * wscale; mode) + residual with the sum taken in the convolution epilogue (the ResBlock
    skip connection, model_spatial_query.py:795-797): no separate pass over the two activations."""

    @staticmethod
    def forward(ctx, x, w, residual, mode, wscale=1.0):
        ctx.save_for_backward(x, w)
        ctx.mode, ctx.wscale = mode, wscale
        return conv_raw(x, pack_weight(w, mode.transposed, wscale, nseg_for(x)), mode, residual=residual)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        mode, wscale = ctx.mode, ctx.wscale
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = TcConv.apply(gy, w, mode.adjoint((x.shape[2], x.shape[3])), wscale)
            if gx.shape[1] != x.shape[1]:
                gx = gx[:, :x.shape[1]]
        if ctx.needs_input_grad[1]:
            gw = TcWeightGrad.apply(x, gy, mode, tuple(w.shape), wscale)
        return gx, gw, (gy if ctx.needs_input_grad[2] else None), None, None


def _fwd_mode(w, stride):
    return Mode("s1" if stride == 1 else "down", w.shape[-1])


def conv2d_bias_act(x, w, bias, stride=1, wscale=1.0, gain=2 ** 0.5):
    """conv2d + bias + leaky_relu(0.2)*gain in one kernel (EqualConv2d + FusedLeakyReLU, StyledConv tail)."""
    if w.shape[-4] % 8:
        raise RuntimeError("conv2d_bias_act needs an output channel count that is a multiple of 8")
    return TcConvBiasAct.apply(x, w, bias, _fwd_mode(w, stride), wscale, gain)


def conv2d_bias_act_carry(x, w, bias, stride=1, wscale=1.0, gain=2 ** 0.5):
    """conv2d_bias_act that returns (out, x_alias); feed x's other consumer from x_alias (see TcConvBiasActCarry)."""
    if w.shape[-4] % 8:
        raise RuntimeError("conv2d_bias_act needs an output channel count that is a multiple of 8")
    return TcConvBiasActCarry.apply(x, w, bias, _fwd_mode(w, stride), wscale, gain)


def conv2d_residual(x, w, residual, stride=1, wscale=1.0):
    """conv2d(x, w * wscale) + residual in one kernel."""
    return TcConvResidual.apply(x, w, residual, _fwd_mode(w, stride), wscale)


def conv2d(x, w, stride=1, wscale=1.0):
    """F.conv2d(x, w * wscale, stride, padding = k//2 if stride == 1 else 0) in bf16 on tensor cores.
    w [O, I, K, K], or [B, O, I, K, K] for per-sample weights (the reference's groups=batch form)."""
    return TcConv.apply(x, w, _fwd_mode(w, stride), wscale)


def conv2d_scaled(x, w, d, stride=1, wscale=1.0):
    """d[b, cout] * conv2d(x, w * wscale): per-(sample, output channel) scale in the convolution epilogue."""
    return TcConvScaled.apply(x, w, d, _fwd_mode(w, stride), wscale)


def conv_transpose2d_scaled(x, w_oi, d, stride=2, wscale=1.0):
    """d[b, cout] * conv_transpose2d(x, (w_oi * wscale)^T, stride 2): the same for the polyphase transposed conv."""
    assert stride == 2
    return TcConvScaled.apply(x, w_oi, d, Mode("up", w_oi.shape[-1]), wscale)


def conv_transpose2d(x, w_oi, stride=2, wscale=1.0):
    """F.conv_transpose2d(x, (w_oi * wscale).transpose(0, 1), stride=2, padding=0), weight kept [O, I, K, K]."""
    assert stride == 2
    return TcConv.apply(x, w_oi, Mode("up", w_oi.shape[-1]), wscale)
