"""CPU restatement of the three `utils.op` operators.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  Nothing under transeditor_b200/ imports it.

Each function cites the reference lines it restates (paths relative to the
reference tree):

* upfirdn2d        utils/op/upfirdn2d.py:143-148 (public signature), :151-186
                   (`upfirdn2d_native`, which is dead code there because it forgets
                   to import F) and the kernel index arithmetic
                   utils/op/upfirdn2d_kernel.cu:71-81,114-129,167-168.
* fused_leaky_relu utils/op/fused_act.py:89-90 and fused_bias_act_kernel.cu:26-47
                   (act=3, grad=0: y = (x+b > 0 ? x+b : alpha*(x+b)) * scale).

Everything is written with differentiable torch ops, so first and second order
gradients of the oracle come from autograd and can be compared with the
hand-written backward kernels of the product.
"""
import torch
import torch.nn.functional as F


def upfirdn2d_planes(x, fir, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """[M, H, W] planes -> [M, H', W'].  Zero-insert, pad/crop, true convolution
    with `fir`, decimate (upfirdn2d.py:151-186; out size upfirdn2d_kernel.cu:167-168)."""
    m, h, w = x.shape
    kh, kw = fir.shape
    # zero insertion: sample (y, x) lands on (y*up_y, x*up_x)
    z = x.new_zeros(m, h, up_y, w, up_x)
    z[:, :, 0, :, 0] = x
    z = z.reshape(m, h * up_y, w * up_x)
    # positive pads add zeros, negative pads crop
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, max(-py0, 0): z.shape[1] - max(-py1, 0), max(-px0, 0): z.shape[2] - max(-px1, 0)]
    # F.conv2d is a cross-correlation; flipping the taps makes it a convolution
    taps = torch.flip(fir, [0, 1]).to(z.dtype).reshape(1, 1, kh, kw)
    full = F.conv2d(z[:, None], taps)[:, 0]
    return full[:, ::down_y, ::down_x]


def upfirdn2d(x, fir, up=1, down=1, pad=(0, 0)):
    """Public form, NCHW in / NCHW out (upfirdn2d.py:143-148: same pad on both axes)."""
    n, c, h, w = x.shape
    out = upfirdn2d_planes(x.reshape(n * c, h, w), fir, up, up, down, down,
                           pad[0], pad[1], pad[0], pad[1])
    return out.reshape(n, c, out.shape[1], out.shape[2])


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """y = scale * leaky_relu(x + bias[c]) with c the dim-1 index (fused_act.py:89-90)."""
    if bias is not None:
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + bias.reshape(shape)
    return F.leaky_relu(x, negative_slope) * scale


class FusedLeakyReLU(torch.nn.Module):
    """Module form with the reference's parameter name (fused_act.py:72-86)."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = torch.nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, x):
        return fused_leaky_relu(x, self.bias, self.negative_slope, self.scale)


def upfirdn2d_index_form(x, fir, up, down, pad0, pad1):
    """Literal numpy-style transcription of the CUDA kernel's per-output index
    arithmetic (upfirdn2d_kernel.cu:84-133) for ONE plane, used to pin the
    formulation above against the kernel's own semantics on tiny inputs."""
    h, w = x.shape
    kh, kw = fir.shape
    oh = (h * up + pad0 + pad1 - kh) // down + 1
    ow = (w * up + pad0 + pad1 - kw) // down + 1
    out = torch.zeros(oh, ow, dtype=x.dtype)
    flipped = torch.flip(fir, [0, 1])
    for oy in range(oh):
        mid_y = oy * down + up - 1 - pad0
        in_y = mid_y // up            # python floor division == floor_div
        tap_y = (in_y + 1) * up - mid_y - 1
        for ox in range(ow):
            mid_x = ox * down + up - 1 - pad0
            in_x = mid_x // up
            tap_x = (in_x + 1) * up - mid_x - 1
            acc = 0.0
            for dy in range(kh // up):
                for dx in range(kw // up):
                    yy, xx = in_y + dy, in_x + dx
                    if 0 <= yy < h and 0 <= xx < w:
                        acc = acc + x[yy, xx] * flipped[tap_y + dy * up, tap_x + dx * up]
            out[oy, ox] = acc
    return out


def linear_interpolate_np(latent_code, boundary, start_distance=-100, end_distance=100, steps=10):
    """numpy restatement of our_interfaceGAN/linear_interpolation.py:36-48 (test oracle for
    transeditor_b200.inference.linear_interpolate): offsets = linspace(start, end, steps); a [1, D] code first
    loses its own projection on the boundary normal; a [1, N, D] code is shifted row-wise by the raw offsets."""
    import numpy as np
    offs = np.linspace(start_distance, end_distance, steps)
    if latent_code.ndim == 2:
        offs = (offs - latent_code.dot(boundary.T)).reshape(-1, 1).astype(np.float32)
        return latent_code + offs * boundary
    offs = offs.reshape(-1, 1, 1).astype(np.float32)
    return latent_code + offs * boundary.reshape(1, 1, -1)
