"""ORL global gather-max-mean timing at the encoder's shapes (CUDA events, L2 flush)."""
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
for B, N, C, k in [(32, 1028, 128, 20), (32, 257, 256, 20), (32, 64, 512, 8)]:
    f = torch.randn(B, N, C, generator=g).cuda()
    idx = torch.randint(0, N, (B, N, k), generator=g, dtype=torch.int32).cuda()
    t = timed(lambda: ops.orl_global(f, idx))
    by = B * (4 * N * C + 4 * N * k)
    print(f"orl_global B={B} N={N} C={C} k={k}: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s ({by/t/1e6/6536*100:.1f}% hbm)")
