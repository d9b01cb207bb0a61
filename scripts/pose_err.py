"""pose-output error of the fused network against the reference golden (tests/golden/posenet.npz), indices replayed."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from util import golden
from tgpose_b200.posenet import PoseNet9D
g = golden("posenet")
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).cuda().eval()
net.face_all.encoder._inject = [torch.from_numpy(g[f"idx_{i:02d}"].astype(np.int32)).cuda() for i in range(14)]
with torch.no_grad():
    torch.manual_seed(7)
    out = net(torch.from_numpy(g["pts"]).cuda(), torch.from_numpy(g["cat_id"]).cuda())
for k in ("p_green_R", "p_red_R"):
    a, b = out[k].cpu().numpy().astype(np.float64), g["out_" + k].astype(np.float64)
    cosang = np.clip((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1)), -1, 1)
    print(k, "deg", np.degrees(np.arccos(cosang)).max())
for k in ("Pred_T", "Pred_s", "f_green_R", "f_red_R"):
    print(k, np.abs(out[k].cpu().numpy() - g["out_" + k]).max())
