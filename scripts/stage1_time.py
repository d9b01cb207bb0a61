"""the heads' first layers (four 1289 -> 1024 convolutions as one 4096-wide contraction, posenet.py stage 1) timed under
CUDA-graph replay in both formulations: over the materialised concatenation (K = 1289), and factored over the upsampling
(per-point K = 265 + coarse products at 257 / 64 points per cloud, added as gathered residuals):
python scripts/stage1_time.py [clouds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tgpose_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, N1, N2 = 1028, 257, 64
M = B * N
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
fine = torch.randn(M, 265, generator=g).cuda()
c1 = torch.randn(B * N1, 512, generator=g).cuda()
c2 = torch.randn(B * N2, 512, generator=g).cuda()
# the points of a cloud come in no particular order, so neither do their nearest coarse points.  STAGE1_SORTED=1: rows sorted by
# their level-1 point, level-2 point mostly shared by neighbouring rows (an ordering that was tried for the heads and dropped:
# it is slower, DESIGN.md 3a')
nn1 = torch.randint(0, N1, (B, N), generator=g)
nn2 = torch.randint(0, N2, (B, N), generator=g)
if os.environ.get("STAGE1_SORTED", "0") == "1":
    nn1 = nn1.sort(dim=1).values
    nn2 = torch.where(torch.rand(B, N, generator=g) < 0.1, nn2, (nn1 * 7) % N2)
nn1, nn2 = nn1.cuda(), nn2.cuda()
cloud = torch.arange(B, device="cuda").view(B, 1)
gi1, gi2 = (nn1 + cloud * N1).int().reshape(-1).contiguous(), (nn2 + cloud * N2).int().reshape(-1).contiguous()
full = torch.cat([fine[:, :256], c1[gi1.long()], c2[gi2.long()], fine[:, 256:]], 1)          # (M, 1289)
W = (torch.randn(4096, 1289, generator=g) * 0.03).cuda()
Wf = torch.cat([W[:, :256], W[:, 1280:]], 1).contiguous()
W1, W2 = W[:, 256:768].contiguous(), W[:, 768:1280].contiguous()
scale, shift = (torch.rand(4096, generator=g) + 0.5).cuda(), torch.randn(4096, generator=g).cuda()
slope = torch.zeros(4096).cuda()
xs_full, xs_f, xs1, xs2 = ops.split_mixed(full), ops.split_mixed(fine), ops.split_mixed(c1), ops.split_mixed(c2)
ws_full, ws_f, ws1, ws2 = ops.split_mixed(W), ops.split_mixed(Wf), ops.split_mixed(W1), ops.split_mixed(W2)


def dests():
    hid = ops.mixed_buf(M, 3072, "cuda")
    mx = torch.full((B, 1024), -2 ** 31, dtype=torch.int32, device="cuda")
    kp = ops.mixed_kpad(3072)
    segs = [(0, 1024, hid, 4, kp), (1024, 2048, hid[:, 512:], 4, kp), (2048, 3072, mx, 3, 0), (3072, 4096, hid[:, 1024:], 4, kp)]
    return hid, mx, segs


hid_a, mx_a, segs_a = dests()
hid_b, mx_b, segs_b = dests()
P1 = torch.empty(B * N1, 4096, device="cuda")
P2 = torch.empty(B * N2, 4096, device="cuda")


def run_full():
    ops.gemm(None, W, True, segs_a, K=1289, A_split=xs_full, B_split=ws_full, mixed=True, scale=scale, shift=shift,
             neg_slope=slope, rows_per_group=N)


def run_coarse():
    ops.gemm(None, W1, True, [(0, 4096, P1, 0, 0)], K=512, A_split=xs1, B_split=ws1, mixed=True)
    ops.gemm(None, W2, True, [(0, 4096, P2, 0, 0)], K=512, A_split=xs2, B_split=ws2, mixed=True)


def run_fine():
    ops.gemm(None, Wf, True, segs_b, K=265, A_split=xs_f, B_split=ws_f, mixed=True, scale=scale, shift=shift,
             neg_slope=slope, rows_per_group=N, res1=P1, res2=P2, res1_idx=gi1, res2_idx=gi2)


def timed(fn, what, flops, out_mb):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{what}: {t * 1e3:.1f} us  ({flops / t / 1e9:.0f} TFLOP/s launched, output {out_mb:.0f} MB -> {out_mb / t / 1e3:.2f} TB/s)")
    return t


out_mb = M * 3072 * 8 / 1e6
timed(run_full, f"stage 1 over the concatenation  M={M} K=1289 N=4096", 2.0 * M * 1289 * 4096, out_mb)
tc = timed(run_coarse, f"coarse products  ({B * N1} + {B * N2}) x 512 x 4096", 2.0 * (B * N1 + B * N2) * 512 * 4096, (B * N1 + B * N2) * 4096 * 4 / 1e6)
tf = timed(run_fine, f"per-point part with gathered residuals  M={M} K=265 N=4096", 2.0 * M * 265 * 4096, out_mb)
print(f"factored total {1e3 * (tc + tf):.1f} us")
kp3 = 3 * ops.mixed_kpad(3072)
a16, b16 = hid_a.view(torch.int16)[:, :kp3], hid_b.view(torch.int16)[:, :kp3]
ha, hb = a16[:, :3072].view(torch.float16).float(), b16[:, :3072].view(torch.float16).float()
print("hidden activations: max abs diff", float((ha - hb).abs().max()), "of scale", float(ha.abs().max()),
      "; pooled max diff", float((ops.decode_max(mx_a) - ops.decode_max(mx_b)).abs().max()))
