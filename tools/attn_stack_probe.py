"""Launch te_attn_stack_fwd / _bwd a few times on the flagship shape (B=16, 8 blocks, 528-wide block 0) — the
target of `ncu -k regex:attn_stack`; also prints CUDA-event times per launch.  Profiling aid, not a bench."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tests.test_gpu_attn_stack import FIELDS, LR, _blocks, _inputs, _to  # noqa: E402
from transeditor_b200 import op  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    tf32 = len(sys.argv) > 3 and sys.argv[3] == "tf32"
    bd = _to(_blocks(8, 528, seed=1), torch.float32, "cuda", grad=True)
    x0, p0, p = [t.float().cuda().requires_grad_(True) for t in _inputs(batch, 528, seed=2)]
    params = [blk[f] for blk in bd for f in FIELDS if blk[f] is not None]
    gy = torch.randn(batch, 16, 512, device="cuda")

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    def fwd():
        with torch.no_grad():
            return op.attn_stack(x0, p0, p, bd, LR, tf32=tf32)

    y = op.attn_stack(x0, p0, p, bd, LR, tf32=tf32)

    def bwd():
        torch.autograd.grad(y, [x0, p0, p] + params, gy, retain_graph=True)

    from transeditor_b200 import lib
    print("resident clusters:", lib.attn_stack_occupancy())
    print("batch %d%s: fwd %.1f us, bwd (data + weight-gradient launches) %.1f us" % (batch, " tf32" if tf32 else "", timed(fwd), timed(bwd)))


if __name__ == "__main__":
    main()
