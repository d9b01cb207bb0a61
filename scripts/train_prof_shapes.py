import sys, os
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from bench import synth_inputs
from tgpose_b200 import _lib
from tgpose_b200.posenet import PoseNet9D
from tgpose_b200.train_step import TrainStep, synthetic_targets
_lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).to(dev)
step = TrainStep(net)
B = 256
pts, cat = synth_inputs(B, 4321)
tgt = synthetic_targets(B, 99, dev)
for _ in range(3):
    step(pts.to(dev), cat.to(dev), tgt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    step(pts.to(dev), cat.to(dev), tgt)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.device_time_total > 300:
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:120]))
rows.sort(reverse=True)
for t, c, k, sh in rows[:28]:
    print(f"{t/1e3:8.3f} ms x{c:3d} {k:22s} {sh}")
