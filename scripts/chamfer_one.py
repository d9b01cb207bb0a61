"""a few chamfer forward launches at BASELINE configs[2] size (for ncu): python scripts/chamfer_one.py [B n m]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
a = [int(v) for v in sys.argv[1:]]
B, n, m = (a + [256, 1028, 1024][len(a):])[:3]
g = torch.Generator().manual_seed(0)
x = torch.rand(B, n, 3, generator=g).cuda()
y = torch.rand(B, m, 3, generator=g).cuda()
d1, d2 = torch.zeros(B, n, device="cuda"), torch.zeros(B, m, device="cuda")
i1, i2 = torch.zeros(B, n, dtype=torch.int32, device="cuda"), torch.zeros(B, m, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.chamfer_forward(x, y, d1, d2, i1, i2)
torch.cuda.synchronize()
print("ok", float(d1.sum()))
