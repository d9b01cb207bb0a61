import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tgpose_b200 import ops, _lib
torch.manual_seed(0)
def run(M, K, N, nk=True):
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') * 0.1 if nk else torch.randn(K, N, device='cuda') * 0.1
    bias = torch.randn(N, device='cuda')
    out = torch.zeros(M, N, device='cuda')
    As = ops.split_tf32(A)
    Bs = ops.split_tf32(W, src_is_kn=not nk)
    ops.gemm(A, W, nk, [(0, N, out, 0, 0)], bias=bias, A_split=As, B_split=Bs)
    torch.cuda.synchronize()
    ref = (A.double() @ (W.double().t() if nk else W.double()) + bias.double())
    err = (out.double() - ref).abs()
    rel = err / (ref.abs() + 1e-3)
    out2 = torch.zeros(M, N, device='cuda')
    ops.gemm(A, W, nk, [(0, N, out2, 0, 0)], bias=bias)
    err2 = (out2.double() - ref).abs()
    print(f"M={M} K={K} N={N} nk={nk}: tc max abs {err.max().item():.3e} max rel {rel.max().item():.3e} | simt max abs {err2.max().item():.3e}", flush=True)
run(128, 32, 64)
run(128, 32, 256)
run(256, 128, 256)
run(1000, 128, 1152, nk=False)
run(32896, 1286, 1024)
run(8224, 256, 200)
# timing
M, K, N = 32896, 1286, 1024
A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda'); out = torch.zeros(M, N, device='cuda')
As = ops.split_tf32(A); Bs = ops.split_tf32(W)
for _ in range(3): ops.gemm(A, W, True, [(0, N, out, 0, 0)], A_split=As, B_split=Bs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.gemm(A, W, True, [(0, N, out, 0, 0)], A_split=As, B_split=Bs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"tc gemm {M}x{K}x{N}: {ms:.3f} ms -> {2*M*K*N/ms/1e9:.1f} TFLOP/s fp32-equivalent ({6*M*K*N/ms/1e9:.1f} TF32 TFLOP/s)")
e0.record()
for _ in range(10): As = ops.split_tf32(A)
e1.record(); torch.cuda.synchronize()
print(f"split A: {e0.elapsed_time(e1)/10:.3f} ms")
