"""one warm-up + one profiled full-network forward (B=32 x 1028) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_inputs
from tgpose_b200.posenet import PoseNet9D
torch.manual_seed(0)
net = PoseNet9D().cuda().eval()
pts, cat = synth_inputs(32, 1234)
pts, cat = pts.cuda(), cat.cuda()
for _ in range(2):
    torch.manual_seed(7)
    with torch.no_grad():
        out = net(pts, cat)
torch.cuda.synchronize()
print("ok", float(out["Pred_T"].sum()))
