"""surface conv timing at the encoder's shape (CUDA events, L2 flush)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def timed(fn, iters=15):
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
B, N, C, k, S = 32, 1028, 128, 20, 7
xyz = torch.rand(B, N, 3, generator=g).cuda()
idx = ops.knn_xyz(xyz, k, want64=False, want32=True)[1]
dirs = ((torch.rand(3, S * C, generator=g) - 0.5) * 0.07).cuda()
t = timed(lambda: ops.surface_conv(xyz, idx, dirs, S, C))
fl = B * (N * k * S * C * 8 + N * S * C)
print(f"surface_conv B={B} N={N} C={C} k={k}: {t*1e3:.1f} us  {fl/t/1e9:.2f} TFLOP/s ({fl/t/1e9/74.4*100:.1f}% fp32)")
