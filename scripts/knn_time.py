"""xyz / feature kNN timing over k (CUDA events, L2 flush): python scripts/knn_time.py [N ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def timed(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
g = torch.Generator().manual_seed(0)
Ns = [int(v) for v in sys.argv[1:]] or [1028, 4096]
for N in Ns:
    B = max(1, 32 * 1028 // N)
    xyz = torch.rand(B, N, 3, generator=g).cuda()
    feat = (torch.randn(B, N, 128, generator=g) * 0.3).cuda()
    for k in (10, 20, 30, 40, 50, 63):
        tx = timed(lambda: ops.knn_xyz(xyz, k, want64=False, want32=True))
        tf = timed(lambda: ops.knn_feat(feat, k, want64=False, want32=True))
        print(f"N={N} B={B} k={k}: xyz {tx*1e3:.1f} us  feat(D=128) {tf*1e3:.1f} us")
