"""Run in a FRESH interpreter by tests/test_gpu_dropin.py: the reference's OWN FaceRecon.py (oracle/_ref/pyref, byte for byte)
with the one-line swap of INTEGRATION.md section 2 -- `network.fs_net_repo.gcn3d` -> `tgpose_b200.gcn3d` -- applied through
sys.modules, i.e. the reference's Face_Enc.__init__ / forward driving our modules through their plain, reference-shaped API
(no extension arguments, int64 indices, torch.cat / BatchNorm / squeeze done by the reference's code).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
sys.path.insert(0, PYREF)

import config.config  # noqa: E402,F401
from absl import flags  # noqa: E402

flags.FLAGS(["dropin"])
import network.fs_net_repo  # noqa: E402  (the reference's package)
import tgpose_b200.gcn3d as ours  # noqa: E402

sys.modules["network.fs_net_repo.gcn3d"] = ours                      # FaceRecon.py:3 now resolves to our module
setattr(sys.modules["network.fs_net_repo"], "gcn3d", ours)
from network.fs_net_repo.FaceRecon import Face_Enc as RefFaceEnc  # noqa: E402
from tgpose_b200.face_enc import Face_Enc as FusedFaceEnc  # noqa: E402

assert RefFaceEnc.__module__ == "network.fs_net_repo.FaceRecon" and "oracle/_ref/pyref" in sys.modules[RefFaceEnc.__module__].__file__

g = np.load(os.path.join(ROOT, "tests", "golden", "face_enc.npz"))
pts, cat = torch.from_numpy(g["pts"]).cuda(), torch.from_numpy(g["cat_id"]).cuda()
torch.manual_seed(0)
ref_on_ours = RefFaceEnc().cuda().eval()                              # the reference's class, our layers inside
torch.manual_seed(0)
fused = FusedFaceEnc().cuda().eval()
sd_ref = {k: v for k, v in ref_on_ours.state_dict().items()}
same_init = all(torch.equal(v, fused.state_dict()[k]) for k, v in sd_ref.items())
names = sorted(sd_ref.keys())
with torch.no_grad():
    torch.manual_seed(7)
    feat_a, fg_a = ref_on_ours(pts, cat)
    torch.manual_seed(7)
    feat_b, _ = fused(pts, cat)


def frac(a, b):
    a, b = a.double().cpu().numpy(), np.asarray(b, np.float64)
    return float((np.abs(a - b) <= 1e-4 * np.maximum(np.abs(a), np.abs(b)) + 1e-6).mean())


print(json.dumps({"shape": list(feat_a.shape), "global_shape": list(fg_a.shape), "same_init": bool(same_init),
                  "state_keys_equal_golden": [str(n) for n in g["param_names"]] == names,
                  "frac_vs_fused": frac(feat_a, feat_b.cpu().numpy()), "frac_vs_reference_golden": frac(feat_a, g["feat"]),
                  "finite": bool(torch.isfinite(feat_a).all())}))
