import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tgpose_b200.posenet import PoseNet9D
from tgpose_b200 import ops
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "posenet.npz"))
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).cuda().eval()
inj = [torch.as_tensor(g[f"idx_{i:02d}"].astype(np.int32)).cuda() for i in range(14)]
net.face_all.encoder._inject = inj
pts, cat = torch.as_tensor(g["pts"]).cuda(), torch.as_tensor(g["cat_id"]).cuda()
with torch.no_grad():
    torch.manual_seed(7)
    out = net(pts, cat)
    feat = out["feat"]
    fcf = feat.permute(0, 2, 1)
    ops.TC_ENABLED = False
    red_simt = net.rot_red(fcf)
    green_simt = net.rot_green(fcf)
    ops.TC_ENABLED = True
    red_tc = net.rot_red(fcf)
    # torch fp64 reference of the red head
    r = net.rot_red.double()
    x = fcf.double()
    import torch.nn.functional as F
    h = F.relu(r.bn1(r.conv1(x))); h = F.relu(r.bn2(r.conv2(h))); h = h.max(2, keepdim=True)[0]
    h = F.relu(r.bn3(r.conv3(h))); red64 = r.conv4(h).squeeze(2)
    net.rot_red.float()
print("red fp64 :", red64.cpu().numpy())
print("red simt :", red_simt.cpu().numpy())
print("red tc   :", red_tc.cpu().numpy())
print("p_red fused:", out["p_red_R"].cpu().numpy(), " golden:", g["out_p_red_R"])
v = red64[:, 1:]; print("p_red fp64 :", (v / (v.norm(dim=1, keepdim=True) + 1e-6)).cpu().numpy())
