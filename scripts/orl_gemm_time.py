"""the five ORL contractions of a forward (conv2(cat[f, g]) + f + f_STE, gcn3d.py:110-112,184-186: W2a.f with two residuals and
the per-cloud term as group bias; single destination, or raw + split destinations where the next layer reads the result as a
tensor-core operand) timed under CUDA-graph replay: python scripts/orl_gemm_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import ops
B = 32
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (N, C, two, post) in [(4, 32, False, False), (64, 64, False, False), (64, 128, False, False), (64, 256, False, False), (1028, 128, True, False), (1028, 128, False, True), (257, 256, True, True), (257, 256, False, True), (64, 512, False, False)]:
    M = B * N
    g = torch.Generator().manual_seed(M + C)
    f = torch.randn(M, C, generator=g).cuda()
    ste = torch.randn(M, C, generator=g).cuda()
    W = (torch.randn(C, C, generator=g) * 0.05).cuda()
    gb = torch.randn(B, C, generator=g).cuda()
    scale, shift = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    out = torch.empty(M, C, device="cuda")
    spl = ops._split_buf(M, C, "cuda")
    segs = [(0, C, out, 0, 0)] + ([(0, C, spl, 2, ops.kpad(C))] if two else [])
    As, Bs = ops.split_tf32(f), ops.split_tf32(W)
    kw = dict(scale=scale, shift=shift, relu=True) if post else {}
    def run():
        ops.gemm(f, W, True, segs, group_bias=gb, rows_per_group=N, res1=f, res2=ste, A_split=As, B_split=Bs, **kw)
    for _ in range(3): run()
    torch.cuda.synchronize()
    v = f.double() @ W.double().t() + f.double() + ste.double() + gb.double().repeat_interleave(N, 0)
    if post:
        v = torch.relu(v * scale.double() + shift.double())
    err = float((out.double() - v).abs().max()) / float(v.abs().max())
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
    torch.cuda.synchronize()
    res = {}
    for warm in (False, True):
        ts = []
        for _ in range(10):
            if not warm:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res[warm] = ts[len(ts) // 2]
    print(f"ORL conv2 M={M} K=N={C} {'raw+split' if two else 'raw      '} {'bn+relu' if post else '       '}: "
          f"{res[False]*1e3:.1f} us L2-flushed, {res[True]*1e3:.1f} us warm  (max rel err {err:.1e})")
