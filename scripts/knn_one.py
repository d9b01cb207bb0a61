"""One xyz kNN launch set (for ncu): 32 x 1028, k = 20."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
g = torch.Generator().manual_seed(0)
x = torch.rand(32, 1028, 3, generator=g).cuda()
for _ in range(3):
    ops.knn_xyz(x, 20, want64=False, want32=True)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    f = (torch.randn(32, 1028, 128, generator=g) * 0.5).cuda()
    for _ in range(3):
        ops.knn_feat(f, 20, want64=False, want32=True)
    torch.cuda.synchronize()
print("ok")
