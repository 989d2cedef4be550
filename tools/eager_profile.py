"""Host-side cost of one EAGER training iteration (the path the reference's unmodified scripts take: no CUDA graphs):
cProfile over plain D step + G step iterations of train_step.Trainer with graphs off.  Profiling aid.

    python tools/eager_profile.py [--precision bf16] > gpurun_out/eager_profile.txt
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from transeditor_b200 import model as te_model
    from transeditor_b200.train_step import TrainConfig, Trainer
    te_model.set_precision(args.precision)
    tr = Trainer(TrainConfig(size=256, batch=16), torch.device("cuda"), seed=0)
    real = torch.rand(16, 3, 256, 256, device="cuda") * 2 - 1
    for _ in range(3):
        tr.iteration = 1
        tr.step(real)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.iteration = 1
        tr.step(real)
    torch.cuda.synchronize()
    print("eager plain iteration: %.1f ms wall" % ((time.perf_counter() - t0) / args.steps * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.steps):
        tr.iteration = 1
        tr.step(real)
    torch.cuda.synchronize()
    pr.disable()
    for key in ("tottime", "cumulative"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
        print(s.getvalue()[:9000])


if __name__ == "__main__":
    main()
