"""Mixed-operand GEMM with the heads' epilogues (mixed-operand store / per-cloud column max): MMA vs epilogue split via
TGP_TC_DEBUG (1: no epilogue, 2: no MMA, 4: no TMA)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps
M, N = 32896, 1028
for K, NC, kind in [(1289, 4096, "mixed"), (1024, 768, "max"), (1289, 512, "mixed"), (512, 512, "mixed")]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(NC, K, device="cuda") * 0.05
    xs, ws = ops.split_mixed(x), ops.split_mixed(w)
    sc = torch.rand(NC, device="cuda") + 0.5; sh = torch.randn(NC, device="cuda"); sl = torch.zeros(NC, device="cuda")
    if kind == "max":
        dst = torch.full((M // N, NC), -2 ** 31, dtype=torch.int32, device="cuda")
        segs = [(0, NC, dst, 3, 0)]
    else:
        dst = ops.mixed_buf(M, NC, "cuda")
        segs = [(0, NC, dst, 4, ops.mixed_kpad(NC))]
    t = ev(lambda: ops.gemm(None, w, True, segs, scale=sc, shift=sh, neg_slope=sl, K=K, A_split=xs, B_split=ws,
                            rows_per_group=N, mixed=True))
    tiles = ((M + 127) // 128) * ((NC + 255) // 256)
    print(f"dbg={os.environ.get('TGP_TC_DEBUG','0')} {kind} {M}x{K}x{NC}: {t*1e3:.1f} us  {2*M*K*NC/t/1e9:.0f} TF/s  per tile-round {t*1e3/(tiles/148):.1f} us", flush=True)
