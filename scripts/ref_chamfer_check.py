"""GPU box: the reference's own chamfer3D kernel (oracle/_ref/chamfer3D, built from /root/reference by oracle/build_ref.py)
against tgp_chamfer_fwd: bit-equality of dist/idx on tie-free inputs and CUDA-event timing of both (same box, same inputs)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tgpose_b200 import ops  # noqa: E402


def main():
    ref = build_ref.load_chamfer()
    assert ref is not None, "oracle/_ref/chamfer3D not built"
    dev = torch.device("cuda:0")
    out = {}
    for (B, n, m) in [(4, 100, 200), (32, 1028, 1024), (256, 1028, 1024)]:
        g = torch.Generator().manual_seed(B * 7 + n)
        a = torch.rand(B, n, 3, generator=g).to(dev)
        b = torch.rand(B, m, 3, generator=g).to(dev)

        def alloc():
            return (torch.zeros(B, n, device=dev), torch.zeros(B, m, device=dev),
                    torch.zeros(B, n, dtype=torch.int32, device=dev), torch.zeros(B, m, dtype=torch.int32, device=dev))
        r = alloc()
        o = alloc()
        ref.forward(a, b, *r)
        ops.chamfer_forward(a, b, *o)
        torch.cuda.synchronize()
        rec = {"dist1_equal": bool(torch.equal(r[0], o[0])), "dist2_equal": bool(torch.equal(r[1], o[1])),
               "idx1_equal": bool(torch.equal(r[2], o[2])), "idx2_equal": bool(torch.equal(r[3], o[3])),
               "dist1_maxdiff": float((r[0] - o[0]).abs().max()), "n_dist1_diff": int((r[0] != o[0]).sum())}

        def timeit(fn, it=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(it):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / it
        # the reference launches on the legacy default stream; torch's current stream is the same one here
        rec["ref_ms"] = timeit(lambda: ref.forward(a, b, *r))
        rec["ours_ms"] = timeit(lambda: ops.chamfer_forward(a, b, *o))
        flops = 2.0 * B * n * m * 9
        rec["ref_frac_fp32"] = flops / (rec["ref_ms"] * 1e-3) / 74.4e12
        rec["ours_frac_fp32"] = flops / (rec["ours_ms"] * 1e-3) / 74.4e12
        out[f"{B}x{n}x{m}"] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
