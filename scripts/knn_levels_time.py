import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tgpose_b200 import _lib, ops
_lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
g = torch.Generator().manual_seed(0)
for (B, N, D, k) in [(32, 1028, 128, 20), (32, 257, 128, 20), (32, 257, 256, 20), (32, 64, 256, 8), (8, 257, 128, 20), (4, 257, 256, 20)]:
    feat = (torch.randn(B, N, D, generator=g) * 0.3).cuda()
    spl = ops.split_tf32(feat.view(B * N, D))
    def fn(): return ops.knn_feat(feat, k, want64=False, want32=True, x_split=spl)[1]
    for _ in range(3): r = fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"B={B} N={N} D={D} k={k}: {ts[5]*1e3:.1f} us  checksum {int(r.long().sum())}")
