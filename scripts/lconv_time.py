"""Times the layer conv kernel at the encoder's four shapes (CUDA events, L2 flush)."""
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
S = 7
for B, N, C, k in [(32, 1028, 128, 20), (32, 257, 256, 20), (32, 64, 512, 8)]:
    xyz = torch.rand(B, N, 3, generator=g).cuda()
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).cuda()
    rec = ops.edge_records(xyz, idx)
    dirs = torch.randn(3, S * C, generator=g).cuda()
    centre = torch.randn(B * N, C, generator=g).cuda()
    slab = torch.randn(C // 4, B * N, S * 4, generator=g).cuda()
    t = timed(lambda: ops.layer_conv(rec, dirs, centre, slab, B, N, S, C))
    fl = B * (N * k * S * C * 9 + N * S * C)
    print(f"layer_conv B={B} N={N} C={C} k={k}: {t*1e3:.1f} us  {fl/t/1e9:.2f} TFLOP/s ({fl/t/1e9/74.4*100:.1f}% fp32)")
