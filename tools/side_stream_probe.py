"""Does the side-stream style work overlap the convolutions?  Times a CUDA-graph replay of G forward+backward
with TE_STYLE_STREAM on/off and reports kernel-time sums per stream from torch.profiler.  Profiling aid."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(flag):
    os.environ["TE_STYLE_STREAM"] = flag
    from transeditor_b200 import model as M
    M.set_precision("bf16")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)
    g = M.Generator(256, 512, 512, 14, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).cuda()
    z = torch.randn(16, 512, 16, device="cuda")
    p = torch.randn(16, 512, 16, device="cuda")

    def step():
        img, _, _ = g(z, p)
        img.square().mean().backward()
        for q in g.parameters():
            q.grad = None

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gr.replay()
        torch.cuda.synchronize()
    per = collections.defaultdict(lambda: [0, 0.0])
    t0, t1 = 1e30, 0
    for ev in prof.events():
        if str(ev.device_type).endswith("CUDA") and ev.device_time > 0:
            st = getattr(ev, "stream", None)
            per[st][0] += 1
            per[st][1] += ev.device_time
            t0 = min(t0, ev.time_range.start); t1 = max(t1, ev.time_range.end)
    print("TE_STYLE_STREAM=%s graph replay %.3f ms; span %.3f ms; per stream (kernels, ms): %s" % (
        flag, ts[len(ts) // 2], (t1 - t0) / 1e3, {k: (v[0], round(v[1] / 1e3, 3)) for k, v in per.items()}))


if __name__ == "__main__":
    run(sys.argv[1])
