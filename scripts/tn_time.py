"""Times the heads' weight-gradient contraction (transposed mixed operands, split-K) against the forward GEMM of the same flops."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import _lib, ops
_lib.load()
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps
M = 263168 // int(os.environ.get("DIV", "1"))
for K1, K2 in [(1289, 4096), (1024, 256), (512, 512)]:
    x = torch.randn(M, K1, device="cuda"); dz = torch.randn(M, K2, device="cuda")
    xt, dzt = ops.split_mixed_t(x), ops.split_mixed_t(dz)
    lib = _lib.load()
    nb = lib.tgp_gemm_tn_tc_workspace(M, K1, K2)
    ws = torch.empty((nb + 3) // 4, dtype=torch.float32, device="cuda")
    out = torch.empty(K1, K2, device="cuda")
    def tn():
        rc = lib.tgp_gemm_tn_tc(xt.data_ptr(), dzt.data_ptr(), M, K1, K2, out.data_ptr(), K2, 1, ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    t = ev(tn)
    fl = 2.0 * M * K1 * K2
    wdummy = torch.randn(K2, K1, device="cuda"); xs, ws2 = ops.split_mixed(x), ops.split_mixed(wdummy)
    z = torch.empty(M, K2, device="cuda")
    t2 = ev(lambda: ops.gemm(None, wdummy, True, [(0, K2, z, 0, 0)], K=K1, A_split=xs, B_split=ws2, mixed=True))
    print(f"M={M} K1={K1} K2={K2}: dW {t:.2f} ms {fl/t/1e9:.0f} TF/s | fwd {t2:.2f} ms {fl/t2/1e9:.0f} TF/s", flush=True)
    del x, dz, xt, dzt, xs, z
