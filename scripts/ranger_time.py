"""Time the clip + Ranger step (tgp_ranger_reduce + tgp_ranger_update) at net1's size against its HBM roofline, next to
torch's clip_grad_norm_ + fused Adam on the same parameters.  Algorithmic bytes per element: reduce 4,
update 28 (36 on Lookahead steps).  The five arenas (5 x 110 MB) exceed the 126 MB L2, so no flush is needed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgpose_b200 import _lib, ops  # noqa: E402
from tgpose_b200.posenet import PoseNet9D  # noqa: E402
from tgpose_b200.ranger import Ranger  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
hbm = float(peaks.get("hbm_gbs", 6536.0))
torch.manual_seed(0)
net = PoseNet9D(train_outputs=True).cuda()
params = [p for p in net.parameters() if p.requires_grad]
n = sum(p.numel() for p in params)
opt = Ranger(params, lr=1e-4)
opt.flat_grads.normal_(0, 1e-3)
ops.EVENT_LOG = None
st = torch.cuda.current_stream()


def timed(fn, iters=24):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ts) / len(ts), ts[len(ts) // 2]


def reduce_only():
    opt._reduce()


def full_step():
    opt.clip_grad_norm_(5.0)
    opt.step()


red_ms, _ = timed(reduce_only)
step_ms, step_med = timed(full_step)           # 24 steps: 4 of them Lookahead steps (k = 6)
upd_ms = step_ms - red_ms
elems = opt._total
res = {"params": n, "arena_elements": elems, "rows": opt.n_rows,
       "reduce_ms": round(red_ms, 4), "reduce_gbs": round(elems * 4 / red_ms / 1e6, 1),
       "update_ms": round(upd_ms, 4), "update_gbs": round(elems * (28 + 8 / 6) / upd_ms / 1e6, 1),
       "step_ms": round(step_ms, 4), "step_median_ms": round(step_med, 4),
       "step_gbs": round(elems * (32 + 8 / 6) / step_ms / 1e6, 1), "hbm_peak_gbs": hbm}
res["step_frac_hbm"] = round(res["step_gbs"] / hbm, 3)
res["update_frac_hbm"] = round(res["update_gbs"] / hbm, 3)

# torch baseline on separate parameter copies: clip_grad_norm_ (foreach) + fused Adam
plain = [torch.nn.Parameter(p.detach().clone()) for p in params]
for p in plain:
    p.grad = torch.randn_like(p) * 1e-3
adam = torch.optim.Adam(plain, lr=1e-4, fused=True)


def torch_step():
    torch.nn.utils.clip_grad_norm_(plain, 5.0)
    adam.step()


res["torch_clip_fused_adam_ms"] = round(timed(torch_step)[0], 4)
print(json.dumps(res))
