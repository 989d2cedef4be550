"""Kernel-time breakdown of the fused pSp encoder (batch 32) with torch.profiler, plus per-head-group timings.
Profiling aid, not a bench:  python tools/encoder_profile.py"""
import sys, os, collections, torch
sys.path.insert(0, os.getcwd())
from torch.profiler import ProfilerActivity, profile
from transeditor_b200.inversion import GradualStyleEncoder, FusedEncoder
torch.manual_seed(0)
enc = GradualStyleEncoder(50, "ir_se").cuda().eval()
fe = FusedEncoder(enc)
x = torch.rand(32, 3, 256, 256, device="cuda") * 2 - 1
for _ in range(3): fe(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fe(x); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if str(ev.device_type).endswith("CUDA") and ev.device_time > 0:
        agg[ev.name][0] += 1; agg[ev.name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print("total %.2f ms, %d kernels" % (tot/1e3, sum(v[0] for v in agg.values())))
for k,(n,us) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print("%9.1f us %5d  %s" % (us, n, k[:140]))
# split: trunk vs heads
import time
def t(fn, it=5):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
c3 = torch.randn(32,512,16,16,device="cuda",dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
p2 = torch.randn(32,512,32,32,device="cuda",dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
p1 = torch.randn(32,512,64,64,device="cuda",dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    print("coarse %.2f middle %.2f fine %.2f spatial %.2f ms" % (t(lambda: fe.coarse(c3)), t(lambda: fe.middle(p2)), t(lambda: fe.fine(p1)), t(lambda: fe.spatial(c3))))
