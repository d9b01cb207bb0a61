"""two warm-up training steps + one profiled step (RL_TDA restatement, B clouds x 1028 points, one GPU) for ncu captures.
Prints `skip=<library launches before the third step> count=<library launches of one step>`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_inputs
from tgpose_b200 import _lib
from tgpose_b200.posenet import PoseNet9D
from tgpose_b200.train_step import TrainStep, augment, synthetic_targets
_lib.load()
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", "64"))
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).to(dev)
net2 = PoseNet9D(only_encoder=True).to(dev)
step = TrainStep(net, net2=net2)
pts, cat = synth_inputs(B, 4321)
aug = augment(pts, 55).to(dev)
tgt = synthetic_targets(B, 99, dev)
pts, cat = pts.to(dev), cat.to(dev)
marks = []
for _ in range(3):
    marks.append(_lib.launch_count())
    loss = step(pts, cat, tgt, aug)
torch.cuda.synchronize()
marks.append(_lib.launch_count())
print(f"skip={marks[2] - marks[0]} count={marks[3] - marks[2]} loss={float(loss):.6f}")
