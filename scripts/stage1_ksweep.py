import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tgpose_b200 import ops
B, N, N1, N2 = 32, 1028, 257, 64
M = B * N
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
nn1 = torch.randint(0, N1, (B, N), generator=g).cuda(); nn2 = torch.randint(0, N2, (B, N), generator=g).cuda()
cloud = torch.arange(B, device="cuda").view(B, 1)
gi1, gi2 = (nn1 + cloud * N1).int().reshape(-1).contiguous(), (nn2 + cloud * N2).int().reshape(-1).contiguous()
P1 = torch.randn(B * N1, 4096, device="cuda"); P2 = torch.randn(B * N2, 4096, device="cuda")
scale, shift = (torch.rand(4096, generator=g) + 0.5).cuda(), torch.randn(4096, generator=g).cuda()
slope = torch.zeros(4096).cuda()
hid = ops.mixed_buf(M, 3072, "cuda"); mx = torch.full((B, 1024), -2 ** 31, dtype=torch.int32, device="cuda")
kp = ops.mixed_kpad(3072)
segs = [(0, 1024, hid, 4, kp), (1024, 2048, hid[:, 512:], 4, kp), (2048, 3072, mx, 3, 0), (3072, 4096, hid[:, 1024:], 4, kp)]
for K in (64, 265):
    fine = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(4096, K, generator=g) * 0.03).cuda()
    xs, ws = ops.split_mixed(fine), ops.split_mixed(W)
    raw = torch.empty(M, 4096, device="cuda")
    for mode in ("res", "nores", "raw", "rawres"):
        kw = dict(res1=P1, res2=P2, res1_idx=gi1, res2_idx=gi2) if mode in ("res", "rawres") else {}
        sg = segs if mode in ("res", "nores") else [(0, 4096, raw, 0, 0)]
        def run():
            ops.gemm(None, W, True, sg, K=K, A_split=xs, B_split=ws, mixed=True, scale=scale, shift=shift, neg_slope=slope, rows_per_group=N, **kw)
        for _ in range(3): run()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        ts.sort()
        print(f"K={K} {mode}: {ts[len(ts)//2]*1e3:.1f} us")
