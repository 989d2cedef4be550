"""Same-GPU baseline: the reference's UNMODIFIED train_spatial_query.py on this B200, (a) with the reference's own
modules + CUDA extensions + cuDNN and (b) with this repository's drop-in modules — both through
tools/run_ref_script.py, stock torch.optim.Adam, eager launches, the script's own loop.

    python tools/reference_gpu_baseline.py [--size 256] [--batch 16] [--out gpurun_out/reference_gpu_baseline.json]

Steady-state iteration time = (wall(iter=K2) - wall(iter=K1)) / (K2 - K1): start-up, the i = 0 sample grid and the
checkpoint write cancel; the window covers the lazy-regulariser cadence (R1 every 16, path every 4)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wall(impl, iters, size, batch, extra):
    with tempfile.TemporaryDirectory() as wd:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "run_ref_script.py"), "--impl", impl, "--workdir", wd] + extra + \
              ["train_spatial_query.py", "synthetic", "--iter", str(iters), "--batch", str(batch), "--size", str(size),
               "--n_sample", "4", "--exp_name", "bench"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("%s failed:\n%s" % (" ".join(cmd), (r.stdout + r.stderr)[-3000:]))
        return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--k1", type=int, default=4)
    ap.add_argument("--k2", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = {}
    runs = [("reference_tf32_off", "reference", ["--tf32", "0"]), ("reference_tf32_on", "reference", ["--tf32", "1"]),
            ("ours_fp32", "ours", ["--precision", "fp32", "--tf32", "0"]), ("ours_bf16", "ours", ["--precision", "bf16", "--tf32", "1"])]
    for key, impl, extra in runs:
        t1 = wall(impl, a.k1, a.size, a.batch, extra)
        t2 = wall(impl, a.k2, a.size, a.batch, extra)
        per_it = (t2 - t1) / (a.k2 - a.k1)
        rows[key] = {"s_per_iteration": round(per_it, 4), "img_per_s": round(a.batch / per_it, 2),
                     "wall_k1_s": round(t1, 1), "wall_k2_s": round(t2, 1)}
        print(key, rows[key], flush=True)
    out = {"what": "unmodified train_spatial_query.py, 1 GPU, eager, torch.optim.Adam, synthetic lmdb", "size": a.size,
           "batch": a.batch, "iterations": [a.k1, a.k2], "runs": rows}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
