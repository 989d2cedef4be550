"""Time each training phase (D step, R1, G step, path regulariser) as a CUDA-graph replay.

    python tools/phase_times.py            # prints ms per phase and the weighted per-iteration average
Profiling aid, not a bench (bench.py reports the real interleaved step).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def profile_phase(tr, real, name, top=45):
    """Eager run of one phase under torch.profiler: device time per kernel name."""
    import collections
    from torch.profiler import ProfilerActivity, profile
    fn = {"d": lambda: tr.d_step(real), "dreg": lambda: tr.d_regularize(real), "g": tr.g_step,
          "greg": tr.g_regularize}[name]
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA") and ev.device_time > 0:
            agg[ev.name][0] += 1
            agg[ev.name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print("phase %s: %d kernels, %.2f ms device time" % (name, sum(v[0] for v in agg.values()), tot / 1e3))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%9.1f us %5d  %s" % (us, n, k[:120]))


def main():
    from transeditor_b200 import model as te_model
    from transeditor_b200.train_step import TrainConfig, Trainer
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    te_model.set_precision("bf16")
    cfg = TrainConfig(size=256, batch=16)
    tr = Trainer(cfg, dev, seed=0)
    real = (torch.rand(cfg.batch, 3, cfg.size, cfg.size) * 2 - 1).to(dev)
    for _ in range(2):
        tr.step(real)
    if len(sys.argv) > 2 and sys.argv[1] == "--profile":
        tr._set_real(real)
        for name in sys.argv[2:]:
            profile_phase(tr, real, name)
        return
    tr.enable_graphs()
    tr.iteration = 0
    tr.step(real)  # captures all four phases
    tr.step(real)
    torch.cuda.synchronize()
    out = {}
    for name in ("d", "dreg", "g", "greg"):
        ga, steps, launches = tr._graphs[name]
        gb = next(iter(steps.values()))[0]
        for which, g in (("fwdbwd", ga), ("optim", gb)):
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            out["%s.%s" % (name, which)] = round(ts[len(ts) // 2], 3)
        out["%s.te_launches" % name] = launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.ema_update()
    e1.record()
    torch.cuda.synchronize()
    out["ema"] = round(e0.elapsed_time(e1) / 5, 3)
    ph = lambda n: out[n + ".fwdbwd"] + out[n + ".optim"]  # noqa: E731
    out["avg_iteration_ms"] = round(ph("d") + ph("g") + ph("dreg") / cfg.d_reg_every + ph("greg") / cfg.g_reg_every
                                    + out["ema"], 3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
