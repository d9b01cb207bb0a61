"""the encoder's four projection launches (centre | slab | f_STE destinations, bias) timed under CUDA-graph replay
(so the host-side tensor-map encodes do not count): python scripts/proj_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import ops
S = 7
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for (M, cin, C) in [(32896, 128, 128), (8224, 128, 256), (8224, 256, 256), (2048, 256, 512)]:
    g = torch.Generator().manual_seed(M + C)
    fm = torch.randn(M, cin, generator=g).cuda()
    ncol = (S + 2) * C
    W = (torch.randn(cin, ncol, generator=g) * 0.05).cuda()
    bias = torch.randn(ncol, generator=g).cuda()
    centre = torch.empty(M, C, device="cuda"); slab = torch.empty(C // 4, M, S * 4, device="cuda"); fste = torch.empty(M, C, device="cuda")
    As, Bs = ops.split_tf32(fm), ops.split_tf32(W, src_is_kn=True)
    segs = [(0, S * C, slab, 1, S * 4), (S * C, S * C + C, centre, 0, 0), (S * C + C, ncol, fste, 0, 0)]
    def run():
        ops.gemm(fm, W, False, segs, bias=bias, A_split=As, B_split=Bs)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ref = fm.double() @ W.double() + bias.double()
    got = torch.cat([slab.permute(1, 0, 2).reshape(M, S * C), centre, fste], 1).double()
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    out_mb = M * ncol * 4 / 1e6
    print(f"proj M={M} K={cin} N={ncol}: {t*1e3:.1f} us  ({2*M*cin*ncol/t/1e9:.0f} TFLOP/s algorithmic, output {out_mb:.0f} MB -> {out_mb/t/1e3:.2f} TB/s)  max rel err {err:.2e}")
