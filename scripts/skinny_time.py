"""Times the per-cloud (M = batch) contractions of the head tails on the exact-fp32 skinny kernel."""
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
for M, K, N in [(32, 5000, 1286), (32, 1024, 5000), (32, 1286, 512), (32, 1024, 1024), (32, 256, 256), (32, 128, 128), (32, 512, 512)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda")
    t = timed(lambda: ops.gemm(A, W, True, [(0, N, out, 0, 0)], tc=False))
    print(f"skinny M={M} K={K} N={N}: {t*1e3:.1f} us  weights {N*K*4/t/1e6:.0f} GB/s")
