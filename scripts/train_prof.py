"""torch.profiler table of one training step (global batch 256 on one GPU): which ATen kernels sit next to ours."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import synth_inputs
from tgpose_b200 import _lib
from tgpose_b200.posenet import PoseNet9D
from tgpose_b200.train_step import TrainStep, augment, synthetic_targets
_lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).to(dev)
net2 = PoseNet9D(only_encoder=True).to(dev)
step = TrainStep(net, net2=net2)
B = int(os.environ.get("B", "256"))
pts, cat = synth_inputs(B, 4321)
tgt = synthetic_targets(B, 99, dev)
aug = augment(pts, 55).to(dev)
for _ in range(3):
    step(pts.to(dev), cat.to(dev), tgt, aug)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(pts.to(dev), cat.to(dev), tgt, aug)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows if "tgp::" in r[0] or "at::" in r[0] or "void" in r[0] or "Memcpy" in r[0] or "Memset" in r[0])
for k, t, c in rows[:45]:
    print(f"{t/1e3:9.3f} ms  x{c:4d}  {k[:110]}")
