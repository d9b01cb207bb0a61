"""Kernel list of ONE CUDA-graph replay of the inference forward (torch.profiler / CUPTI): durations and gaps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import synth_inputs
from tgpose_b200 import _lib
from tgpose_b200.posenet import PoseNet9D
from tgpose_b200.graph import GraphedPoseNet
_lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = PoseNet9D().to(dev).eval()
pts, cat = synth_inputs(32, 1234)
pts, cat = pts.to(dev), cat.to(dev)
g = GraphedPoseNet(net, 32, 1028)
for _ in range(5):
    torch.manual_seed(7); g(pts, cat)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    torch.manual_seed(7); g(pts, cat)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
t1 = max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"kernels {len(evs)}  span {(t1-t0)/1e3:.3f} ms  sum of durations {busy/1e3:.3f} ms")
# union of busy intervals (several streams)
cur_s, cur_e, union = None, None, 0
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None: union += cur_e - cur_s
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
union += cur_e - cur_s
print(f"union busy {union/1e3:.3f} ms  idle {(t1-t0-union)/1e3:.3f} ms")
agg = {}
for e in evs:
    k = e.name.split("(")[0][-60:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{t/1e3:8.3f} ms x{c:3d}  {k}")
if os.environ.get("TIMELINE"):
    print("--- timeline: start(us) dur(us) idle-before(us) name  (idle = no kernel on any stream)")
    busy_end = evs[0].time_range.start
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        gap = max(0, s - busy_end)
        print(f"{(s-t0):8.1f} {(en-s):7.1f} {gap:6.1f}  {e.name.split('(')[0][-48:]}")
        busy_end = max(busy_end, en)
