"""Times the xyz / feature kNN kernels (CUDA events, L2 flush between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def timed(fn, iters=20):
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
for B, N, k in [(32, 1028, 20), (32, 1028, 4), (32, 257, 20), (32, 64, 8), (32, 1028, 30), (8, 4096, 20), (2, 16384, 20)]:
    x = torch.rand(B, N, 3, generator=g).cuda()
    t = timed(lambda: ops.knn_xyz(x, k, want64=False, want32=True))
    print(f"xyz  B={B} N={N} k={k}: {t*1e3:.1f} us  {B*N*N/t/1e6:.0f} Gpairs/s")
for B, N, D, k in [(32, 1028, 128, 20), (32, 257, 128, 20), (32, 257, 256, 20), (32, 64, 256, 8), (32, 1028, 128, 30), (8, 4096, 128, 20)]:
    x = (torch.randn(B, N, D, generator=g) * 0.5).cuda()
    t = timed(lambda: ops.knn_feat(x, k, want64=False, want32=True))
    print(f"feat B={B} N={N} D={D} k={k}: {t*1e3:.1f} us  {B*N*N/t/1e6:.0f} Gpairs/s")
