"""Same-GPU baseline of the STEP itself: the reference's own classes + its own CUDA extensions + cuDNN + torch.optim.Adam
(oracle/train_cpu.RefClassCpuTrainer on cuda: the loop body of train_spatial_query.py:166-306 without the script's
loader / logging), device-resident synthetic batch, CUDA-event timed, lazy-regulariser cadence included.

    python tools/reference_gpu_step.py [--steps 32] > gpurun_out/reference_gpu_step.json
Needs an unmodified reference tree (baseline/_ref)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    from oracle.train_cpu import RefClassCpuTrainer
    out = {"what": "reference classes + reference CUDA ops + cuDNN, eager, torch.optim.Adam, 256^2, batch %d, %d steps "
                   "(R1 every 16, path every 4)" % (a.batch, a.steps)}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        tr = RefClassCpuTrainer(size=256, batch=a.batch, device="cuda")
        real = torch.rand(a.batch, 3, 256, 256, device="cuda") * 2 - 1
        for _ in range(3):
            tr.it = 1
            tr.step(real)
        tr.it = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            tr.step(real)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out["tf32_%s" % ("on" if tf32 else "off")] = {"ms_per_step": round(ms, 2), "img_per_s": round(a.batch / ms * 1e3, 2)}
        del tr
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
