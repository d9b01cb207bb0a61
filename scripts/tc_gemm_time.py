import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import ops
def t(M, K, N, reps=10):
    A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda'); out = torch.zeros(M, N, device='cuda')
    As = ops.split_tf32(A); Bs = ops.split_tf32(W)
    for _ in range(3): ops.gemm(A, W, True, [(0, N, out, 0, 0)], A_split=As, B_split=Bs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.gemm(A, W, True, [(0, N, out, 0, 0)], A_split=As, B_split=Bs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"dbg={os.environ.get('TGP_TC_DEBUG','0')} {M}x{K}x{N}: {ms:.3f} ms  {6*M*K*N/ms/1e9:.1f} TF32 TFLOP/s", flush=True)
if os.environ.get("TGP_SHAPES"):          # e.g. TGP_SHAPES=32896x128x1024,8224x256x2048
    for spec in os.environ["TGP_SHAPES"].split(","):
        t(*[int(v) for v in spec.split("x")])
    sys.exit(0)
t(32896, 1286, 1024)
t(32896, 128, 1152)
t(32896, 1024, 256)
if os.environ.get("TGP_SMALL"):
    t(32896, 128, 128)
    t(8224, 256, 256)
    t(2048, 512, 512)
