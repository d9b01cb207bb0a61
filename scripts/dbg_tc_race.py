import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import ops
torch.manual_seed(1)
def check(M, K, N, nseg_split=True, reps=6):
    A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda') * 0.05
    sc = torch.rand(N, device='cuda') + 0.5; sh = torch.randn(N, device='cuda'); sl = torch.zeros(N, device='cuda')
    As = ops.split_tf32(A); Ws = ops.split_tf32(W)
    ref = torch.empty(M, N, device='cuda')
    ops.gemm(A, W, True, [(0, N, ref, 0, 0)], scale=sc, shift=sh, neg_slope=sl, tc=False)
    outs = []
    for r in range(reps):
        q = N // 4
        kp = ops.kpad(q)
        bufs = [ops._split_buf(M, q, 'cuda') for _ in range(3)]
        raw = torch.empty(M, q, device='cuda')
        segs = [(0, q, bufs[0], 2, kp), (q, 2 * q, bufs[1], 2, kp), (2 * q, 3 * q, raw, 0, 0), (3 * q, 4 * q, bufs[2], 2, kp)]
        ops.gemm(None, W, True, segs, scale=sc, shift=sh, neg_slope=sl, K=K, A_split=As, B_split=Ws)
        full = torch.cat([bufs[0][:, :q] + bufs[0][:, kp:kp + q], bufs[1][:, :q] + bufs[1][:, kp:kp + q], raw,
                          bufs[2][:, :q] + bufs[2][:, kp:kp + q]], 1)
        outs.append(full)
    torch.cuda.synchronize()
    errs = [(o - ref).abs().max().item() for o in outs]
    same = [torch.equal(outs[0], o) for o in outs[1:]]
    print(f"M={M} K={K} N={N}: max err vs simt per rep {['%.2e' % e for e in errs]} bit-identical reps {same}", flush=True)
check(256, 1289, 4096)
check(256, 1024, 1024)
check(32896, 1289, 4096, reps=3)
check(256, 512, 2048)
check(1000, 128, 512)
