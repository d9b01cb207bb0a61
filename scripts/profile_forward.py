"""two warm-up forwards + one profiled full-network forward (B=32 x 1028, eval, eager launches) for ncu captures.
Prints `skip=<library launches before the third forward> count=<library launches of one forward>`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_inputs
from tgpose_b200 import _lib
from tgpose_b200.posenet import PoseNet9D
torch.manual_seed(0)
net = PoseNet9D().cuda().eval()
pts, cat = synth_inputs(32, 1234)
pts, cat = pts.cuda(), cat.cuda()
marks = []
for _ in range(3):
    marks.append(_lib.launch_count())
    torch.manual_seed(7)
    with torch.no_grad():
        out = net(pts, cat)
torch.cuda.synchronize()
marks.append(_lib.launch_count())
print(f"skip={marks[2] - marks[0]} count={marks[3] - marks[2]} checksum={float(out['Pred_T'].sum()):.6f}")
