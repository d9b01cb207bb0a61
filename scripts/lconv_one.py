"""one layer-conv launch at the conv_1 shape (for ncu): python scripts/lconv_one.py [N C k B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
a = [int(v) for v in sys.argv[1:]]
N, C, k, B = (a + [1028, 128, 20, 32][len(a):])[:4]
S = 7
g = torch.Generator().manual_seed(0)
xyz = torch.rand(B, N, 3, generator=g).cuda()
idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).cuda()
rec = ops.edge_records(xyz, idx)
dirs = torch.randn(3, S * C, generator=g).cuda()
centre = torch.randn(B * N, C, generator=g).cuda()
slab = torch.randn(C // 4, B * N, S * 4, generator=g).cuda()
for _ in range(3):
    out = ops.layer_conv(rec, dirs, centre, slab, B, N, S, C)
torch.cuda.synchronize()
print("ok", float(out.sum()))
