"""ctypes binding of libte_b200.so (the C ABI declared in include/te_b200.h).

There is NO fallback: if the shared library is missing or a tensor is not on a CUDA
device, these functions raise.  Tensors are passed as raw device pointers; kernels are
enqueued on torch's current CUDA stream; outputs are allocated by the caller with
torch.empty (PyTorch is plumbing here: memory, streams, autograd bookkeeping).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libte_b200.so")

TE_F32, TE_BF16, TE_F16, TE_F64 = 0, 1, 2, 3
_DTYPES = {torch.float32: TE_F32, torch.bfloat16: TE_BF16, torch.float16: TE_F16,
           torch.float64: TE_F64}


class ConvGeom(ctypes.Structure):
    """Mirror of te_conv_geom."""
    _fields_ = [("batch", ctypes.c_int), ("cin", ctypes.c_int), ("hin", ctypes.c_int),
                ("win", ctypes.c_int), ("cout", ctypes.c_int), ("hout", ctypes.c_int),
                ("wout", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
                ("up", ctypes.c_int), ("down", ctypes.c_int), ("pad_y", ctypes.c_int),
                ("pad_x", ctypes.c_int), ("flip", ctypes.c_int),
                ("w_so", ctypes.c_int64), ("w_si", ctypes.c_int64),
                ("act", ctypes.c_int), ("noise_bstride", ctypes.c_int64)]


class TcConvDesc(ctypes.Structure):
    """Mirror of te_tc_conv_desc."""
    _fields_ = [("batch", ctypes.c_int), ("hin", ctypes.c_int), ("win", ctypes.c_int), ("cin", ctypes.c_int),
                ("hout", ctypes.c_int), ("wout", ctypes.c_int), ("cout", ctypes.c_int),
                ("ntaps", ctypes.c_int), ("tap_dy", ctypes.c_int * 9), ("tap_dx", ctypes.c_int * 9),
                ("tap_w", ctypes.c_int * 9), ("w_slices", ctypes.c_int),
                ("in_stride", ctypes.c_int), ("out_stride", ctypes.c_int),
                ("out_off_y", ctypes.c_int), ("out_off_x", ctypes.c_int),
                ("grid_h", ctypes.c_int), ("grid_w", ctypes.c_int),
                ("act", ctypes.c_int), ("out_f32", ctypes.c_int), ("w_bstride", ctypes.c_int64),
                ("act_gain", ctypes.c_float), ("wgrad_alpha", ctypes.c_float), ("residual", ctypes.c_void_p),
                ("slope", ctypes.c_void_p), ("split", ctypes.c_int), ("reserved", ctypes.c_int)]


ATTN_FIELDS = ("w_proj", "b_proj", "w_q", "b_q", "w_k", "b_k", "w_v", "b_v", "w_o", "b_o", "w_m1", "b_m1",
               "w_m2", "b_m2")


class AttnBlock(ctypes.Structure):
    """Mirror of te_attn_block."""
    _fields_ = [(f, ctypes.c_void_p) for f in ATTN_FIELDS] + [("in_dim", ctypes.c_int), ("param_dim", ctypes.c_int)]


class PackTask(ctypes.Structure):
    """Mirror of te_pack_task."""
    _fields_ = [("src", ctypes.c_void_p), ("dst_n", ctypes.c_void_p), ("dst_t", ctypes.c_void_p),
                ("out_ch", ctypes.c_int), ("in_ch", ctypes.c_int), ("taps", ctypes.c_int), ("scale", ctypes.c_float),
                ("split", ctypes.c_int)]


class LinearTask(ctypes.Structure):
    """Mirror of te_linear_task."""
    _fields_ = [("x", ctypes.c_void_p), ("x_rs", ctypes.c_int64), ("x_cs", ctypes.c_int64),
                ("w", ctypes.c_void_p), ("w_ld", ctypes.c_int64), ("w_trans", ctypes.c_int),
                ("bias", ctypes.c_void_p), ("bias_mul", ctypes.c_float),
                ("y", ctypes.c_void_p), ("y_rs", ctypes.c_int64), ("y_cs", ctypes.c_int64),
                ("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int), ("alpha", ctypes.c_float),
                ("act", ctypes.c_int), ("pixel_norm", ctypes.c_int), ("k_splits", ctypes.c_int),
                ("rnorm_out", ctypes.c_void_p)]


class LinearWgradTask(ctypes.Structure):
    """Mirror of te_linear_wgrad_task."""
    _fields_ = [("g", ctypes.c_void_p), ("g_rs", ctypes.c_int64), ("g_cs", ctypes.c_int64),
                ("x", ctypes.c_void_p), ("x_rs", ctypes.c_int64), ("x_cs", ctypes.c_int64),
                ("x_scale", ctypes.c_void_p), ("gw", ctypes.c_void_p), ("gbias", ctypes.c_void_p),
                ("m", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int),
                ("alpha", ctypes.c_float), ("bias_mul", ctypes.c_float)]


_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_SIGNATURES = {
    "te_version": ([], _I),
    "te_last_error": ([], ctypes.c_char_p),
    "te_fused_bias_act": ([_P, _P, _P, _P, _I, _I, _F, _F, _L, _L, _L, _I, _P], _I),
    "te_fused_bias_act_bwd": ([_P, _P, _P, _P, _F, _F, _L, _L, _L, _I, _P], _I),
    "te_upfirdn2d": ([_P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "te_upfirdn2d_bias_act": ([_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P], _I),
    "te_conv2d_simt": ([_P, _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(ConvGeom), _I, _P], _I),
    "te_conv2d_wgrad_simt": ([_P, _P, _P, _P, _P, ctypes.POINTER(ConvGeom), _I, _P], _I),
    "te_adam_ema": ([_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _F, _P], _I),
    "te_adam_ema_devstep": ([_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _P, _F, _F, _P], _I),
    "te_conv_tc": ([_P, _P, _P, _P, _P, ctypes.POINTER(TcConvDesc), _P], _I),
    "te_conv_wgrad_tc": ([_P, _P, _P, ctypes.POINTER(TcConvDesc), _P], _I),
    "te_wgrad_unpack": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "te_weight_energy": ([_P, _P, _L, _I, _F, _P], _I),
    "te_weight_energy_bwd": ([_P, _P, _P, _L, _I, _F, _P], _I),
    "te_scale_bc": ([_P, _P, _P, _L, _L, _I, _I, _P], _I),
    "te_dot_bc": ([_P, _P, _P, _L, _L, _I, _I, _P], _I),
    "te_split_bf16": ([_P, _P, _P, _L, _L, _I, _I, _P], _I),
    "te_conv2d_tc": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _P], _I),
    "te_gemm_tc_selftest": ([_P, _P, _P, _I, _I, _I, _P], _I),
    "te_attn_core": ([_P, _P, _P, _P, _P, _I, _I, _P], _I),
    "te_pack_weights_tc": ([ctypes.POINTER(PackTask), _I, _P], _I),
    "te_linear_grouped": ([ctypes.POINTER(LinearTask), _I, _I, _P], _I),
    "te_linear_wgrad_grouped": ([ctypes.POINTER(LinearWgradTask), _I, _P], _I),
    "te_from_rgb_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _I, _P], _I),
    "te_from_rgb_bwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _I, _P], _I),
    "te_image_prep": ([_P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "te_image_quantize": ([_P, _P, _I, _I, _I, _L, _L, _L, _L, _F, _F, _I, _P], _I),
    "te_attn_stack_workspace": ([ctypes.POINTER(AttnBlock), _I, _I, ctypes.POINTER(_L), ctypes.POINTER(_L)], _I),
    "te_attn_stack_occupancy": ([ctypes.POINTER(_I), ctypes.POINTER(_I)], _I),
    "te_attn_stack_fwd": ([_P, _P, _P, _P, ctypes.POINTER(AttnBlock), _I, _I, _F, _I, _P, _P], _I),
    "te_attn_stack_bwd": ([_P, _P, _P, ctypes.POINTER(AttnBlock), _P, _P, _P, _P, ctypes.POINTER(AttnBlock), _I, _I,
                           _F, _I, _P, _P, _P], _I),
}

_lib = None
emulated = False   # set by the CPU test-suite's emulation fixture only (tests/emu.py); the product never sets it
launch_count = 0  # number of te_* kernel launches issued through this binding (bench.py reads it)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "transeditor_b200: %s is missing — run `python -m transeditor_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the build is stale
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().te_last_error().decode("utf-8", "replace")
        raise RuntimeError("transeditor_b200.%s failed (status %d): %s" % (what, rc, msg))


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError("transeditor_b200: unsupported dtype %s" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("transeditor_b200: expected a CUDA tensor, got device '%s' "
                               "(the hot path has no CPU fallback)" % t.device)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _count(n=1):
    global launch_count
    launch_count += n


# ----------------------------------------------------------------------------- thin wrappers
TE_BIAS_F32 = 0x100


def fused_bias_act(out, x, bias, ref, act, grad, alpha, scale, step_b, size_b):
    code = dtype_code(x)
    if bias is not None and bias.dtype == torch.float32 and x.dtype in (torch.bfloat16, torch.float16):
        code |= TE_BIAS_F32
    _check(load().te_fused_bias_act(ptr(out), ptr(x), ptr(bias), ptr(ref), act, grad, alpha, scale,
                                    x.numel(), step_b, size_b, code, stream()),
           "fused_bias_act")
    _count()


def fused_bias_act_bwd(grad_in, grad_bias, g, ref, alpha, scale, step_b, size_b):
    _check(load().te_fused_bias_act_bwd(ptr(grad_in), ptr(grad_bias), ptr(g), ptr(ref), alpha, scale,
                                        g.numel(), step_b, size_b, dtype_code(g), stream()),
           "fused_bias_act_bwd")
    _count()


def upfirdn2d(out, x, fir, major, in_h, in_w, minor, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    kh, kw = fir.shape
    _check(load().te_upfirdn2d(ptr(out), ptr(x), ptr(fir), major, in_h, in_w, minor, kh, kw, up_x, up_y,
                               down_x, down_y, px0, px1, py0, py1, dtype_code(x), stream()),
           "upfirdn2d")
    _count()


def conv2d_simt(y, x, w, in_scale, out_scale, bias, noise, noise_w, geom):
    _check(load().te_conv2d_simt(ptr(y), ptr(x), ptr(w), ptr(in_scale), ptr(out_scale), ptr(bias),
                                 ptr(noise), ptr(noise_w), ctypes.byref(geom), dtype_code(x), stream()),
           "conv2d_simt")
    _count()


def conv2d_wgrad_simt(gw, x, gy, in_scale, out_scale, geom):
    _check(load().te_conv2d_wgrad_simt(ptr(gw), ptr(x), ptr(gy), ptr(in_scale), ptr(out_scale),
                                       ctypes.byref(geom), dtype_code(x), stream()),
           "conv2d_wgrad_simt")
    _count()


def adam_ema(p, g, m, v, ema, lr, beta1, beta2, eps, step, ema_decay, grad_scale):
    _check(load().te_adam_ema(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), p.numel(), lr, beta1, beta2,
                              eps, step, ema_decay, grad_scale, stream()), "adam_ema")
    _count()


def adam_ema_devstep(p, g, m, v, ema, lr, beta1, beta2, eps, step_dev, ema_decay, grad_scale):
    _check(load().te_adam_ema_devstep(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), p.numel(), lr, beta1, beta2,
                                      eps, ptr(step_dev), ema_decay, grad_scale, stream()), "adam_ema_devstep")
    _count()


def attn_core(out, sim, q, k, v, batch, tokens):
    _check(load().te_attn_core(ptr(out), ptr(sim), ptr(q), ptr(k), ptr(v), batch, tokens, stream()),
           "attn_core")
    _count()


def pack_weights_tc(tasks):
    """tasks: list of (src f32 [O, I, K, K] contiguous, dst_n bf16 or None, dst_t bf16 or None, scale[, nseg]);
    nseg = 2 or 3 writes that many split-operand planes ([nseg, K*K, ., .] destinations)."""
    if not tasks:
        return
    table = (PackTask * len(tasks))()
    for e, task in zip(table, tasks):
        src, dst_n, dst_t, scale = task[:4]
        e.split = task[4] if len(task) > 4 else 1
        if src.dtype != torch.float32 or not src.is_contiguous() or src.dim() != 4 or src.shape[2] != src.shape[3]:
            raise TypeError("pack_weights_tc: source must be a contiguous float32 [O, I, K, K] tensor")
        for d in (dst_n, dst_t):
            if d is not None and (d.dtype != torch.bfloat16 or not d.is_contiguous()):
                raise TypeError("pack_weights_tc: destinations must be contiguous bfloat16 tensors")
        e.src = src.data_ptr()
        e.dst_n = None if dst_n is None else dst_n.data_ptr()
        e.dst_t = None if dst_t is None else dst_t.data_ptr()
        e.out_ch, e.in_ch, e.taps, e.scale = src.shape[0], src.shape[1], src.shape[2] * src.shape[3], scale
    _check(load().te_pack_weights_tc(table, len(tasks), stream()), "pack_weights_tc")
    _count((len(tasks) + 63) // 64)


def _attn_table(blocks):
    """blocks: list of dicts {field: tensor or None for each of ATTN_FIELDS, "in_dim": int, "param_dim": int}
    (parameters, or the buffers receiving their gradients) -> C array of te_attn_block."""
    table = (AttnBlock * len(blocks))()
    for entry, blk in zip(table, blocks):
        for f in ATTN_FIELDS:
            t = blk.get(f)
            if t is not None:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise TypeError("attn_stack: %s must be a contiguous float32 tensor" % f)
                setattr(entry, f, t.data_ptr())
        entry.in_dim, entry.param_dim = blk["in_dim"], blk["param_dim"]
    return table


def attn_stack_workspace(blocks, batch):
    """(save_floats, gws_floats) for te_attn_stack_fwd / te_attn_stack_bwd."""
    table = _attn_table(blocks)
    a, b = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(load().te_attn_stack_workspace(table, len(blocks), batch, ctypes.byref(a), ctypes.byref(b)),
           "attn_stack_workspace")
    return a.value, b.value


def attn_stack_occupancy():
    """Samples resident at once on the current device: {"fwd": (8-CTA clusters, 4-CTA clusters), "bwd": (...)}."""
    a, b = (ctypes.c_int * 2)(), (ctypes.c_int * 2)()
    _check(load().te_attn_stack_occupancy(a, b), "attn_stack_occupancy")
    return {"fwd": (a[0], a[1]), "bwd": (b[0], b[1])}


def attn_stack_fwd(y, x0, p0, p, blocks, batch, lr_mul, tf32, save):
    table = _attn_table(blocks)
    _check(load().te_attn_stack_fwd(ptr(y), ptr(x0), ptr(p0), ptr(p), table, len(blocks), batch, lr_mul,
                                    1 if tf32 else 0, ptr(save), stream()), "attn_stack_fwd")
    _count()


def attn_stack_bwd(g_x0, g_p0, g_p, grads, gy, x0, p0, p, blocks, batch, lr_mul, tf32, save, gws):
    table, gtable = _attn_table(blocks), _attn_table(grads)
    _check(load().te_attn_stack_bwd(ptr(g_x0), ptr(g_p0), ptr(g_p), gtable, ptr(gy), ptr(x0), ptr(p0), ptr(p), table,
                                    len(blocks), batch, lr_mul, 1 if tf32 else 0, ptr(save), ptr(gws), stream()),
           "attn_stack_bwd")
    _count(2)


def conv2d_tc(y, x, w, out_scale, bias, batch, hin, win, cin, cout, kh, kw, act, w_bstride):
    _check(load().te_conv2d_tc(ptr(y), ptr(x), ptr(w), ptr(out_scale), ptr(bias), batch, hin, win, cin,
                               cout, kh, kw, act, w_bstride, stream()), "conv2d_tc")
    _count()


def conv_tc(y, x, w, out_scale, bias, desc):
    _check(load().te_conv_tc(ptr(y), ptr(x), ptr(w), ptr(out_scale), ptr(bias), ctypes.byref(desc), stream()),
           "conv_tc")
    _count()


def conv_wgrad_tc(gw, g, x, desc):
    _check(load().te_conv_wgrad_tc(ptr(gw), ptr(g), ptr(x), ctypes.byref(desc), stream()), "conv_wgrad_tc")
    _count()


def scale_bc(y, x, s, batch, pixels, channels):
    _check(load().te_scale_bc(ptr(y), ptr(x), ptr(s), batch, pixels, channels, dtype_code(x), stream()), "scale_bc")
    _count()


def split_bf16(dst, x, s, batch, pixels, channels, nseg):
    _check(load().te_split_bf16(ptr(dst), ptr(x), ptr(s), batch, pixels, channels, nseg, stream()), "split_bf16")
    _count()


def dot_bc(out, a, b, batch, pixels, channels):
    _check(load().te_dot_bc(ptr(out), ptr(a), ptr(b), batch, pixels, channels, dtype_code(a), stream()), "dot_bc")
    _count()


def gemm_tc_selftest(d, a, b, m, n, k):
    _check(load().te_gemm_tc_selftest(ptr(d), ptr(a), ptr(b), m, n, k, stream()), "gemm_tc_selftest")
    _count()


def image_prep(dst_nchw, dst_nhwc8, src_hwc, flip, batch, h, w):
    code = TE_F32 if dst_nhwc8 is None else dtype_code(dst_nhwc8)
    _check(load().te_image_prep(ptr(dst_nchw), ptr(dst_nhwc8), ptr(src_hwc), ptr(flip), batch, h, w, code, stream()),
           "image_prep")
    _count()


def image_quantize(dst_hwc, src, low, high):
    b, _, h, w = src.shape
    sb, sc, sy, sx = src.stride()
    _check(load().te_image_quantize(ptr(dst_hwc), ptr(src), b, h, w, sb, sc, sy, sx, low, high, dtype_code(src),
                                    stream()), "image_quantize")
    _count()


def _f32_2d(t, what):
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError("grouped linear: %s must be a 2-D float32 tensor, got %s %s" % (what, t.dtype, tuple(t.shape)))


def linear_grouped(tasks, tf32=False):
    """tasks: list of dicts {x [M,K] (any strides), w ([N,K] contiguous; or [K,N] contiguous with w_trans=True),
    y [M,N] (any strides), bias [N] or None, bias_mul, alpha, act (0/1), pixel_norm (bool), rnorm_out [M] or None,
    k_splits}.  One launch per 32 tasks (te_linear_grouped)."""
    if not tasks:
        return
    table = (LinearTask * len(tasks))()
    for e, t in zip(table, tasks):
        x, w, y, bias = t["x"], t["w"], t["y"], t.get("bias")
        require_cuda(x, w, y, bias, t.get("rnorm_out"))
        _f32_2d(x, "x"), _f32_2d(w, "w"), _f32_2d(y, "y")
        if not w.is_contiguous() or (bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous())):
            raise TypeError("grouped linear: weights and biases must be contiguous float32 tensors")
        trans = bool(t.get("w_trans", False))
        m, k = x.shape
        n = w.shape[1] if trans else w.shape[0]
        if (w.shape[0] if trans else w.shape[1]) != k or tuple(y.shape) != (m, n) or (bias is not None and bias.numel() != n):
            raise RuntimeError("grouped linear: shapes x %s, w %s (trans=%s), y %s do not agree"
                               % (tuple(x.shape), tuple(w.shape), trans, tuple(y.shape)))
        e.x, e.x_rs, e.x_cs = x.data_ptr(), x.stride(0), x.stride(1)
        e.w, e.w_ld, e.w_trans = w.data_ptr(), w.shape[1], int(trans)
        e.bias = None if bias is None else bias.data_ptr()
        e.bias_mul = float(t.get("bias_mul", 1.0))
        e.y, e.y_rs, e.y_cs = y.data_ptr(), y.stride(0), y.stride(1)
        e.m, e.n, e.k, e.alpha = m, n, k, float(t.get("alpha", 1.0))
        e.act, e.pixel_norm, e.k_splits = int(t.get("act", 0)), int(bool(t.get("pixel_norm", False))), int(t.get("k_splits", 1))
        rn = t.get("rnorm_out")
        e.rnorm_out = None if rn is None else rn.data_ptr()
    _check(load().te_linear_grouped(table, len(tasks), 1 if tf32 else 0, stream()), "linear_grouped")
    _count((len(tasks) + 31) // 32)


def linear_wgrad_grouped(tasks):
    """tasks: list of dicts {g [M,N], x [M,K] (any strides), gw [N,K] contiguous, gbias [N] or None, x_scale [M] or
    None, alpha, bias_mul}.  One launch per 32 tasks (te_linear_wgrad_grouped)."""
    if not tasks:
        return
    table = (LinearWgradTask * len(tasks))()
    for e, t in zip(table, tasks):
        g, x, gw, gb, xs = t["g"], t["x"], t["gw"], t.get("gbias"), t.get("x_scale")
        require_cuda(g, x, gw, gb, xs)
        _f32_2d(g, "g"), _f32_2d(x, "x"), _f32_2d(gw, "gw")
        m, n = g.shape
        k = x.shape[1]
        if x.shape[0] != m or tuple(gw.shape) != (n, k) or not gw.is_contiguous():
            raise RuntimeError("grouped linear wgrad: shapes g %s, x %s, gw %s do not agree"
                               % (tuple(g.shape), tuple(x.shape), tuple(gw.shape)))
        e.g, e.g_rs, e.g_cs = g.data_ptr(), g.stride(0), g.stride(1)
        e.x, e.x_rs, e.x_cs = x.data_ptr(), x.stride(0), x.stride(1)
        e.x_scale = None if xs is None else xs.data_ptr()
        e.gw = gw.data_ptr()
        e.gbias = None if gb is None else gb.data_ptr()
        e.m, e.n, e.k = m, n, k
        e.alpha, e.bias_mul = float(t.get("alpha", 1.0)), float(t.get("bias_mul", 1.0))
    _check(load().te_linear_wgrad_grouped(table, len(tasks), stream()), "linear_wgrad_grouped")
    _count((len(tasks) + 31) // 32)


def from_rgb_fwd(y, x, w, bias, wscale, slope, gain):
    b, _, h, wd = x.shape
    _check(load().te_from_rgb_fwd(ptr(y), ptr(x), ptr(w), ptr(bias), b, h, wd, w.shape[0], wscale, slope, gain,
                                  dtype_code(y), stream()), "from_rgb_fwd")
    _count()


def from_rgb_bwd(gw, gbias, gx, g, out, x, w, wscale, slope, gain):
    b, _, h, wd = x.shape
    _check(load().te_from_rgb_bwd(ptr(gw), ptr(gbias), ptr(gx), ptr(g), ptr(out), ptr(x), ptr(w), b, h, wd, w.shape[0],
                                  wscale, slope, gain, dtype_code(g), stream()), "from_rgb_bwd")
    _count()


def wgrad_unpack(out, ws, batch, o_dim, i_dim, taps, rows, ld, trans, clear=False):
    _check(load().te_wgrad_unpack(ptr(out), ptr(ws), batch, o_dim, i_dim, taps, rows, ld, int(trans), int(clear),
                                  stream()), "wgrad_unpack")
    _count()


TE_ERR_UNSUPPORTED = -2


def upfirdn2d_bias_act(out, x, fir, bias, major, in_h, in_w, minor, px0, px1, py0, py1, slope, gain):
    """Returns False (nothing launched) when the geometry is not covered by the fused kernel."""
    kh, kw = fir.shape
    rc = load().te_upfirdn2d_bias_act(ptr(out), ptr(x), ptr(fir), ptr(bias), major, in_h, in_w, minor, kh, kw, px0, px1,
                                      py0, py1, slope, gain, dtype_code(x), stream())
    if rc == TE_ERR_UNSUPPORTED:
        return False
    _check(rc, "upfirdn2d_bias_act")
    _count()
    return True


def weight_energy(energy, w, rows, taps, coef):
    _check(load().te_weight_energy(ptr(energy), ptr(w), rows, taps, coef, stream()), "weight_energy")
    _count()


def weight_energy_bwd(gw, w, g, rows, taps, coef2):
    _check(load().te_weight_energy_bwd(ptr(gw), ptr(w), ptr(g), rows, taps, coef2, stream()), "weight_energy_bwd")
    _count()
