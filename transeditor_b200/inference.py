"""Latency path for editing / interpolation.

The reference scripts walk a latent code across a semantic boundary on the HOST (numpy), then push every step
through `g_ema` one frame at a time: numpy -> torch -> GPU -> PIL per frame
(`our_interfaceGAN/edit_all_noinversion_ffhq.py:235-280`, `test_spatial_query.py:96-105`,
`our_interfaceGAN/linear_interpolation.py:4-48`).  A batch-1 generator forward is ~1 500 kernel launches of a few
microseconds each, so the frame time is launch latency.  Here:

* `linear_interpolate`  — same signature and semantics as the reference's numpy helper, on whatever device the
  code lives on (no host round trip);
* `GraphedGenerator`    — `Generator.forward` for fixed shapes / flags captured once into a CUDA graph and replayed
  per frame;
* `edit_frames`         — the inner loop of the editing scripts: interpolate Z+ and/or P+ and render every step,
  post-processed to uint8 on the GPU.
"""
import numpy as np
import torch


def linear_interpolate(latent_code, boundary, start_distance=-100, end_distance=100, steps=10):
    """Move `latent_code` along the unit normal `boundary`: `steps` codes whose signed distance to the boundary's
    hyperplane runs linearly from start_distance to end_distance (`linear_interpolation.py:4-48`).

    latent_code [1, D]      -> [steps, D]; the code's own projection on the normal is removed first, so the
                               distances are absolute;
    latent_code [1, N, D]   -> [steps, N, D]; every one of the N rows is shifted by the same offsets (W+ / Z+ /
                               P+ style codes), distances relative to the code.
    Accepts numpy arrays (returns numpy, float32 offsets like the reference) or torch tensors (stays on device)."""
    is_np = isinstance(latent_code, np.ndarray)
    code = torch.as_tensor(latent_code)
    normal = torch.as_tensor(boundary).to(code.device)
    if not (code.shape[0] == 1 and normal.dim() == 2 and normal.shape[0] == 1 and normal.shape[1] == code.shape[-1]):
        raise AssertionError("linear_interpolate: need latent_code [1, D] or [1, N, D] and boundary [1, D]")
    offsets = torch.linspace(float(start_distance), float(end_distance), int(steps), dtype=torch.float64,
                             device=code.device)
    if code.dim() == 2:
        offsets = offsets - (code.to(torch.float64) @ normal.to(torch.float64).t()).reshape(-1)
        out = code + offsets.reshape(-1, 1).to(torch.float32) * normal
    elif code.dim() == 3:
        out = code + offsets.reshape(-1, 1, 1).to(torch.float32) * normal.reshape(1, 1, -1)
    else:
        raise ValueError("Input `latent_code` should be with shape [1, latent_space_dim] or "
                         "[1, N, latent_space_dim]; got %s" % (tuple(code.shape),))
    return out.numpy() if is_np else out


class GraphedGenerator:
    """`generator(style, op_param, **flags)` under no_grad for FIXED input shapes, replayed from a CUDA graph.

        gg = GraphedGenerator(g_ema, batch=1, use_style_mapping=False, use_spatial_mapping=False)
        img = gg(z_plus, p_plus)          # [1, 3, S, S] f32, valid until the next call

    The first call runs three eager warm-up forwards (lazy initialisation), captures the graph and replays it;
    every later call costs two small copies into the static input buffers plus one graph launch.  Outputs are
    static buffers owned by the graph: `.clone()` what must outlive the next call.  Weights are read at replay
    time (the captured graph contains the f32 -> bf16 weight repack kernels), so in-place updates of the generator
    (EMA, load_state_dict) are picked up without re-capturing — provided no tc.PackCache is installed while the
    graph is captured (cached weight copies would be frozen into it; a Trainer's cache is only active inside the
    trainer's own phases, see tc.use_pack_cache)."""

    def __init__(self, generator, batch, **flags):
        self.generator = generator
        self.batch = batch
        self.flags = dict(flags)
        self._graph = None
        self._in = None
        self._out = None

    def _capture(self, style, op_param):
        self._in = (style.detach().clone(), op_param.detach().clone())
        with torch.no_grad():
            for _ in range(3):
                self.generator(*self._in, **self.flags)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.generator(*self._in, **self.flags)
        self._graph, self._out = graph, out

    def __call__(self, style, op_param):
        if not (style.is_cuda and op_param.is_cuda):
            raise RuntimeError("GraphedGenerator needs CUDA inputs")
        if style.shape[0] != self.batch or op_param.shape[0] != self.batch:
            raise RuntimeError("GraphedGenerator was built for batch %d" % self.batch)
        if self._graph is None:
            self._capture(style, op_param)
        elif style.shape != self._in[0].shape or op_param.shape != self._in[1].shape:
            raise RuntimeError("GraphedGenerator: input shapes changed since capture (%s, %s)" %
                               (tuple(self._in[0].shape), tuple(self._in[1].shape)))
        self._in[0].copy_(style, non_blocking=True)
        self._in[1].copy_(op_param, non_blocking=True)
        self._graph.replay()
        return self._out


def to_uint8(images, bgr=False):
    """[-1, 1] float NCHW -> uint8 NHWC on the device (the `clamp.add(1).div(2).mul(255).round()` of the editing
    scripts, `edit_all_noinversion_ffhq.py:246-247`, plus the channel-last permute of `make_image`)."""
    if bgr:
        images = images[:, [2, 1, 0]]
    return images.clamp(-1, 1).add(1).mul(127.5).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def edit_frames(graphed, z_plus, p_plus, z_boundary=None, p_boundary=None, z_distance=0.0, p_distance=0.0, steps=10):
    """One edited sequence: `steps` frames walking Z+ and/or P+ from -distance to +distance across their
    boundaries, rendered with the mapping networks bypassed (the loops at
    `edit_all_noinversion_ffhq.py:235-280`).  z_plus / p_plus are [1, 16, 512] (token-major, as the scripts store
    them); boundaries [1, 512] or None to keep that code fixed.  `graphed` is a batch-1 GraphedGenerator built with
    use_style_mapping=False, use_spatial_mapping=False.  Returns uint8 [steps, S, S, 3] on the device."""
    def walk(code, boundary, dist):
        if boundary is None:
            return code.expand(steps, *code.shape[1:])
        return linear_interpolate(code, boundary.to(code.device), start_distance=-dist, end_distance=dist, steps=steps)

    zs = walk(z_plus, z_boundary, z_distance)
    ps = walk(p_plus, p_boundary, p_distance)
    frames = []
    for j in range(steps):
        out = graphed(zs[j:j + 1].transpose(1, 2), ps[j:j + 1].transpose(1, 2))
        img = out[0] if isinstance(out, (tuple, list)) else out
        frames.append(to_uint8(img))
    return torch.cat(frames)
