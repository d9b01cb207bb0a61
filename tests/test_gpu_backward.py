"""GPU parity tests of the backward kernels (north_star item 5, SURVEY 8a'): the CUDA path through the C-ABI
against (1) the numpy backward oracle on seeded inputs and (2) gradients produced by the unmodified reference's
own autograd (tests/golden/backward.npz, face_enc_bwd.npz; made by make_golden.py: backward_cases).

Tolerance (util.assert_grad_close): |a-b| <= 1e-4*max(|a|,|b|) + 2e-5*max|ref| per tensor."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from util import assert_close, assert_grad_close, golden

pytestmark = pytest.mark.gpu


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def nump(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from tgpose_b200 import _lib, ops as o
    _lib.load()
    return o


def _params(g, prefix):
    return {k[len(prefix):]: torch.as_tensor(g[k]) for k in g.files if k.startswith(prefix)}


# ----------------------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("M,K1,K2,tc", [(1000, 128, 1152, True), (5000, 16, 48, True), (257, 3, 70, False),
                                        (4096, 128, 3, False), (2, 32, 32, False), (40000, 256, 256, True),
                                        (33, 130, 7, False)])
def test_gemm_tn(ops, M, K1, K2, tc):
    g = torch.Generator().manual_seed(M + K1 + K2)
    A = torch.randn(M, K1, generator=g)
    B = torch.randn(M, K2, generator=g)
    out = nump(ops.gemm_tn(A.cuda(), B.cuda(), tc=tc))
    ref = A.double().t() @ B.double()
    # fp32 accumulation over M terms of unit scale: error ~ sqrt(M) * 2^-24 * |terms|
    assert np.abs(out - ref.numpy()).max() <= 1e-4 * max(1.0, float(ref.abs().max())) * 0.2
    # strided operands / strided output (column blocks of the projection's gradient operand)
    big = torch.randn(M, K1 + K2 + 5, generator=g).cuda()
    dst = torch.zeros(K1, K2 + 9).cuda()
    ops.gemm_tn(big[:, :K1], big[:, K1 + 5:], out=dst[:, 4:4 + K2], tc=tc)
    ref2 = big[:, :K1].double().t() @ big[:, K1 + 5:].double()
    assert np.abs(nump(dst[:, 4:4 + K2]) - nump(ref2)).max() <= 2e-5 * max(1.0, float(ref2.abs().max()))
    assert float(dst[:, :4].abs().max()) == 0.0 and float(dst[:, 4 + K2:].abs().max()) == 0.0


@pytest.mark.parametrize("M,K1,K2", [(1000, 128, 1152), (40000, 256, 1289), (5001, 70, 48), (33000, 1024, 256)])
def test_gemm_tn_mixed(ops, M, K1, K2):
    """weight-gradient contraction of the heads on transposed MIXED operands (fp16 hi.hi + bf16 cross terms, split-K):
    error at the fp32-summation-noise level of the output scale, deterministic, strided operands accepted."""
    g = torch.Generator().manual_seed(M + K1 + K2)
    A = torch.randn(M, K1, generator=g).cuda()
    B = torch.randn(M, K2 + 3, generator=g).cuda()[:, 3:]
    out = ops.gemm_tn(A, B, tc=True, mixed=True)
    ref = A.double().t() @ B.double()
    assert float((out.double() - ref).abs().max()) <= 1.6e-5 * float(ref.abs().max()) * max(1.0, (M / 4096) ** 0.5)
    assert torch.equal(out, ops.gemm_tn(A, B, tc=True, mixed=True))


def test_gemm_tn_tc_is_deterministic(ops):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(30000, 128, generator=g).cuda()
    B = torch.randn(30000, 256, generator=g).cuda()
    a = nump(ops.gemm_tn(A, B, tc=True))
    b = nump(ops.gemm_tn(A, B, tc=True))
    assert np.array_equal(a, b)          # split-K partials are reduced in a fixed order


@pytest.mark.parametrize("M,C,rpg", [(2056, 128, 1028), (1000, 37, None), (64, 512, 8), (263168, 64, None)])
def test_colsum(ops, M, C, rpg):
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C + 3, generator=g).cuda()
    out = nump(ops.colsum(x[:, 1:1 + C], rows_per_group=rpg))
    r = M if rpg is None else rpg
    ref = x[:, 1:1 + C].double().view(M // r, r, C).sum(1)
    assert np.abs(out - nump(ref)).max() <= 1e-5 * np.sqrt(r) * 4


def test_act_bwd(ops):
    g = torch.Generator().manual_seed(3)
    grad, y = torch.randn(300, 70, generator=g).cuda(), torch.randn(300, 70, generator=g).cuda()
    sc = torch.rand(70, generator=g).cuda() + 0.5
    out = ops.act_bwd(grad, y, sc, True)
    assert torch.equal(out, grad * (y > 0) * sc)
    dst = torch.zeros(300, 100).cuda()
    ops.act_bwd(grad, None, None, False, out=dst[:, 10:80])
    assert torch.equal(dst[:, 10:80], grad) and float(dst[:, :10].abs().max()) == 0


@pytest.mark.parametrize("C,rows", [(128, False), (24, True), (7, False)])
def test_gather_max_bwd(ops, C, rows):
    g = torch.Generator().manual_seed(C)
    B, N, k = 3, 200, 9
    f = torch.randn(B, N, C, generator=g)
    x = torch.rand(B, N, 3, generator=g)
    idx = orc.knn_xyz(x.numpy(), k)
    r = torch.randperm(N, generator=g)[:50] if rows else None
    out, arg = ops.gather_max(f.cuda(), cu(idx), rows=r.cuda() if rows else None, want_arg=True)
    G = torch.randn(out.shape, generator=g)
    df = torch.zeros(B, N, C).cuda()
    ops.gather_max_bwd(G.cuda(), cu(idx), arg, N, df, rows=r.cuda() if rows else None)
    ref = orc.gather_max_backward(f.numpy(), idx, G.numpy(), rows=r.numpy() if rows else None)
    assert_grad_close(nump(df), ref, what="gather_max_bwd")
    # per-cloud broadcast gradient with the 1/N of the ORL mean
    if not rows:
        dg = torch.randn(B, C, generator=g)
        df2 = torch.zeros(B, N, C).cuda()
        ops.gather_max_bwd(dg.cuda(), cu(idx), arg, N, df2, per_cloud=True, scale=1.0 / N)
        ref2 = orc.gather_max_backward(f.numpy(), idx, np.broadcast_to((dg.numpy() / N)[:, None, :], (B, N, C)))
        assert_grad_close(nump(df2), ref2, what="orl scatter")


def test_scatter_add_rows(ops):
    g = torch.Generator().manual_seed(11)
    B, N, M, k, C = 2, 64, 257, 1, 48
    idx = torch.randint(0, N, (B, M, k), generator=g)
    G = torch.randn(B, M, k, C, generator=g)
    out = nump(ops.scatter_add_rows(G.cuda(), idx.cuda(), N))
    ref = torch.zeros(B, N, C, dtype=torch.float64)
    for b in range(B):
        ref[b].index_add_(0, idx[b].reshape(-1), G[b].reshape(-1, C).double())
    assert_grad_close(out, ref.numpy(), what="scatter_add_rows")


@pytest.mark.parametrize("S,C,N,k,B", [(7, 128, 1028, 20, 2), (7, 32, 100, 9, 3), (4, 8, 50, 33, 1), (7, 256, 257, 20, 2)])
def test_layer_conv_bwd_vs_oracle(ops, S, C, N, k, B):
    g = torch.Generator().manual_seed(S * C + N)
    x = torch.rand(B, N, 3, generator=g)
    fmk = torch.randn(B, N, 16, generator=g)
    idx = orc.knn_feat(fmk.numpy(), k)
    dirs = torch.rand(3, S * C, generator=g) - 0.5
    P = torch.randn(B, N, (S + 1) * C, generator=g)
    G = torch.randn(B, N, C, generator=g)
    M = B * N
    # forward through the CUDA path (slab layout) to get the saved arg-max slots
    slab = P[..., C:].reshape(M, S, C // 4, 4).permute(2, 0, 1, 3).contiguous().cuda()
    centre = P[..., :C].reshape(M, C).contiguous().cuda()
    rec = ops.edge_records(x.cuda(), cu(idx, torch.int32))
    out, arg = ops.layer_conv(rec, dirs.cuda(), centre, slab, B, N, S, C, want_arg=True)
    assert_close(nump(out), orc.layer_conv(x.numpy(), idx, dirs.numpy(), P.numpy(), S, C), what="layer conv fwd")
    dP = torch.full((M, (S + 2) * C + 8), 7.0).cuda()
    dsup = dP[:, C:C + S * C]
    dd = ops.layer_conv_bwd(rec, dirs.cuda(), slab, arg, G.view(M, C).cuda(), B, N, S, C, d_support=dsup)
    ref_dP, ref_dd = orc.layer_conv_backward(x.numpy(), idx, dirs.numpy(), P.numpy(), S, C, G.numpy())
    mine = nump(dsup).reshape(M, C // 4, S, 4).transpose(0, 2, 1, 3).reshape(B, N, S * C)   # slab order -> (s, c)
    assert_grad_close(mine, ref_dP[..., C:], what="d_support")
    assert_grad_close(nump(dd), ref_dd, what="d_directions (layer)")
    assert float((dP[:, :C] - 7).abs().max()) == 0 and float((dP[:, C + S * C:] - 7).abs().max()) == 0


@pytest.mark.parametrize("S,C,N,k", [(7, 128, 1028, 20), (7, 32, 100, 9), (3, 20, 64, 5)])
def test_surface_conv_bwd_vs_oracle(ops, S, C, N, k):
    g = torch.Generator().manual_seed(S + C + N)
    B = 2
    x = torch.rand(B, N, 3, generator=g)
    idx = orc.knn_xyz(x.numpy(), k)
    dirs = torch.rand(3, S * C, generator=g) - 0.5
    G = torch.randn(B, N, C, generator=g)
    out, arg = ops.surface_conv(x.cuda(), cu(idx, torch.int32), dirs.cuda(), S, C, want_arg=True)
    dd = ops.surface_conv_bwd(x.cuda(), cu(idx, torch.int32), dirs.cuda(), arg, G.view(B * N, C).cuda(), S, C)
    ref = orc.surface_conv_backward(x.numpy(), idx, dirs.numpy(), S, C, G.numpy())
    assert_grad_close(nump(dd), ref, what="d_directions (surface)")


# ----------------------------------------------------------------------------------------- modules vs reference autograd
def _load(mod, sd):
    mod.load_state_dict(sd)
    return mod.cuda().train()


def test_hs_surface_module_backward_golden():
    from tgpose_b200 import gcn3d
    g = golden("backward")
    m = _load(gcn3d.HSlayer_surface(16, 7), _params(g, "s_p_"))
    out = m(cu(g["x"]), int(g["k"]))
    assert_close(nump(out), g["s_out"], what="surface fwd (train)")
    (out * cu(g["s_G"])).sum().backward()
    for n, p in m.named_parameters():
        assert_grad_close(nump(p.grad), g["s_g_" + n], what=f"HSlayer_surface grad {n}")
    # ... and against the numpy oracle
    idx = g["idx_xyz"].astype(np.int64)
    og = orc.hs_surface_backward({k: v.numpy() for k, v in _params(g, "s_p_").items()}, g["x"], int(g["k"]), idx, idx, g["s_G"])
    for n, p in m.named_parameters():
        assert_grad_close(nump(p.grad), og[n], what=f"HSlayer_surface grad {n} vs oracle")


def test_hs_layer_module_backward_golden():
    from tgpose_b200 import gcn3d
    g = golden("backward")
    m = _load(gcn3d.HS_layer(16, 32, 7), _params(g, "l_p_"))
    fm = cu(g["l_fm"]).requires_grad_(True)
    out = m(cu(g["x"]), fm, int(g["k"]), idx_feat=cu(g["l_idx"], torch.int32))
    assert_close(nump(out), g["l_out"], what="layer fwd (train)")
    (out * cu(g["l_G"])).sum().backward()
    assert_grad_close(nump(fm.grad), g["l_dfm"], what="HS_layer d feature_map")
    for n, p in m.named_parameters():
        assert_grad_close(nump(p.grad), g["l_g_" + n], what=f"HS_layer grad {n}")


def test_pool_module_backward_golden():
    from tgpose_b200 import gcn3d
    g = golden("backward")
    pool = gcn3d.Pool_layer(4, 4)
    f = cu(g["p_f"]).requires_grad_(True)
    torch.manual_seed(7)
    _, fp = pool(cu(g["x"]), f)
    (fp * cu(g["p_G"])).sum().backward()
    assert_grad_close(nump(f.grad), g["p_df"], what="Pool_layer d feature_map")


def test_face_enc_backward_golden():
    """Face_Enc in train mode (BatchNorm batch statistics) with the reference's 14 index tensors replayed:
    every parameter gradient of sum(feat * W) against the reference's autograd."""
    from tgpose_b200.face_enc import Face_Enc
    g = golden("face_enc_bwd")
    torch.manual_seed(0)
    enc = Face_Enc().cuda().train()
    gen = torch.Generator().manual_seed(1234)
    B, N = 2, 128
    pts = torch.rand(B, N, 3, generator=gen)
    cat_id = torch.randint(0, 6, (B, 1), generator=gen).float()
    pts = pts - pts.mean(dim=1, keepdim=True)
    assert np.array_equal(pts.numpy(), g["pts"])
    W = torch.randn(B, N, 1286, generator=gen) * 0.1
    enc._inject = [cu(g[f"idx_{i:02d}"].astype(np.int32)) for i in range(14)]
    torch.manual_seed(7)
    feat, _ = enc(pts.cuda(), cat_id.cuda())
    # train-mode BatchNorm subtracts the batch mean: entries that are small residues of O(1) values carry the
    # absolute error of the O(1) values, so the floor is 2e-5 (x the unit feature scale) instead of 1e-6
    assert_close(nump(feat).reshape(-1)[::97], g["feat_sample"], floor=2e-5, what="Face_Enc train-mode feat")
    (feat * W.cuda()).sum().backward()
    checked = 0
    worst = 0.0
    for n, p in enc.named_parameters():
        if n.startswith("proj_layer"):
            continue
        assert p.grad is not None, n
        gnp = nump(p.grad).reshape(-1)
        ref_sum = g["gsum_" + n]
        # sampled entries: tolerance relative to the mean |grad| scale of the tensor
        scale = ref_sum[1] / gnp.size
        sel, val = g["gsel_" + n], g["gval_" + n]
        err = np.abs(gnp[sel].astype(np.float64) - val)
        tol = 1e-3 * np.maximum(np.abs(val), np.abs(gnp[sel])) + 1e-3 * scale
        # max-over-neighbours is discontinuous: where two candidates are equal within fp32 rounding, the reference and
        # this path may pick different winners and route one contribution to a different row (same situation as the
        # feature-space kNN rule, SURVEY 8c').  Gate: >= 99.5 % of the sampled entries inside the tolerance, none
        # further off than 5 % of the tensor's mean |grad|.
        frac_bad = float((err > tol).mean())
        worst = max(worst, frac_bad)
        assert frac_bad <= 5e-3 and err.max() <= 0.05 * max(scale, float(np.abs(val).max())), \
            f"{n}: {(err > tol).sum()}/{err.size} sampled gradient entries off; worst {err.max():.3e} (scale {scale:.3e})"
        assert abs(np.abs(gnp.astype(np.float64)).sum() - ref_sum[1]) <= 1e-3 * ref_sum[1], n
        checked += 1
    print(f"Face_Enc backward: {checked} parameter tensors checked, worst out-of-tolerance fraction {worst:.2e}")
    assert checked == 29    # conv_0: 3, conv_1..4: 5 each, bn1..3: 2 each


def test_full_size_backward_properties():
    """BASELINE-size clouds (1028 points): finite gradients for every encoder parameter, gradient of a sum over a
    batch equals the sum of per-shard gradients (linearity over clouds; BatchNorm in eval mode so clouds are independent)."""
    from tgpose_b200.face_enc import Face_Enc
    torch.manual_seed(0)
    enc = Face_Enc().cuda().eval()
    gen = torch.Generator().manual_seed(9)
    # 8 clouds, shards of 4: every level keeps >= 256 rows per call, so the same (tensor-core) kernels run in all
    # three calls and a cloud's forward is bit-identical in each (DESIGN.md "Determinism")
    pts = torch.rand(8, 1028, 3, generator=gen).cuda()
    cat = torch.randint(0, 6, (8, 1), generator=gen).float().cuda()
    W = torch.randn(8, 1028, 1286, generator=gen).cuda() * 0.01
    names = [n for n, _ in enc.named_parameters() if not n.startswith("proj_layer") and not n.startswith("bn")]

    def grads(lo, hi):
        enc.zero_grad(set_to_none=True)
        torch.manual_seed(7)
        feat, _ = enc(pts[lo:hi].contiguous(), cat[lo:hi].contiguous())
        (feat * W[lo:hi]).sum().backward()
        return {n: p.grad.detach().clone() for n, p in enc.named_parameters() if n in names}

    full = grads(0, 8)
    a, b = grads(0, 4), grads(4, 8)
    for n in names:
        assert torch.isfinite(full[n]).all(), n
        s = a[n] + b[n]
        scale = float(full[n].abs().max())
        assert float((full[n] - s).abs().max()) <= 1e-4 * scale + 1e-7, (n, float((full[n] - s).abs().max()), scale)


# ----------------------------------------------------------------------------------------- loss tail + training step
@pytest.mark.parametrize("B,n,m", [(4, 1028, 1024), (2, 100, 200), (1, 7, 3)])
def test_dcd_vs_oracle(B, n, m):
    from tgpose_b200.dist_chamfer_3D import calc_dcd
    g = torch.Generator().manual_seed(n + m)
    a = (torch.rand(B, n, 3, generator=g) * 0.3).cuda().requires_grad_(True)
    b = (torch.rand(B, m, 3, generator=g) * 0.3).cuda()
    loss, d1, d2, i1, i2 = calc_dcd(a, b, alpha=70, n_lambda=0.3, return_raw=True)
    ref = orc.calc_dcd(nump(d1), nump(d2), nump(i1), nump(i2), alpha=70.0, n_lambda=0.3)
    assert_close(nump(loss), ref, what="calc_dcd")
    # gradient: finite differences of the oracle loss are discontinuous in idx, so compare with torch autograd
    # through the reference's own formula evaluated on our dist/idx (weights detached, TDA_loss_sym_recon.py:433)
    loss.sum().backward()
    d1r = d1.detach().clone().requires_grad_(True)
    l1 = []
    for bb in range(B):
        c1 = torch.bincount(i1[bb].long(), minlength=m)
        w1 = (c1[i1[bb].long()].float() ** 0.3 + 1e-6) ** (-1) * (m / n)
        l1.append((-torch.exp(-d1r[bb] * 70) * w1 + 1.).mean())
    torch.stack(l1).sum().backward()
    # chain through chamfer backward on both sides: compare d loss / d dist1 via the saved coefficients instead
    from tgpose_b200 import ops
    _, c1, _ = ops.dcd(d1.detach(), d2.detach(), i1, i2, 70.0, 0.3)
    assert_close(nump(c1), nump(d1r.grad), rel=1e-4, floor=1e-7, what="d dcd / d dist1")
    assert torch.isfinite(a.grad).all()


@pytest.mark.parametrize("tag", list("abcd"))
def test_dcd_vs_reference_golden(tag):
    """calc_dcd / calc_cd and their gradient w.r.t. the prediction against the REFERENCE'S OWN functions
    (losses/TDA_loss_sym_recon.py:411-450, :495-509 executed from source by make_golden.py with the reference's pure-torch
    chamfer): loss rel 1e-4, gradient by the gradient rule of tests/util.py."""
    from tgpose_b200.dist_chamfer_3D import calc_cd, calc_dcd
    from util import golden
    g = golden("dcd")
    alpha, lam, non_reg = float(g[f"{tag}_kw"][0]), float(g[f"{tag}_kw"][1]), bool(g[f"{tag}_kw"][2])
    pred = torch.from_numpy(g[f"{tag}_pred"]).cuda().requires_grad_(True)
    gt = torch.from_numpy(g[f"{tag}_gt"]).cuda()
    loss = calc_dcd(pred, gt, alpha=alpha, n_lambda=lam, non_reg=non_reg)
    cd_p, cd_t = calc_cd(pred, gt)
    assert_close(nump(loss), g[f"{tag}_loss"], what=f"calc_dcd {tag}")
    assert_close(nump(cd_p), g[f"{tag}_cd_p"], what="cd_p")
    assert_close(nump(cd_t), g[f"{tag}_cd_t"], what="cd_t")
    (loss.sum() + 0.25 * cd_t.sum()).backward()
    assert_grad_close(nump(pred.grad), g[f"{tag}_grad"], what=f"d (dcd + cd_t) / d pred {tag}")


@pytest.mark.parametrize("B,n,m", [(3, 300, 100), (2, 64, 257)])
def test_dcd_non_reg_vs_oracle(B, n, m):
    """calc_dcd(non_reg=True): both point-count ratios clamped to >= 1 (TDA_loss_sym_recon.py:418-420)."""
    from tgpose_b200.dist_chamfer_3D import calc_dcd
    g = torch.Generator().manual_seed(n * 7 + m)
    a = (torch.rand(B, n, 3, generator=g) * 0.3).cuda()
    b = (torch.rand(B, m, 3, generator=g) * 0.3).cuda()
    loss, d1, d2, i1, i2 = calc_dcd(a, b, alpha=40, n_lambda=0.5, return_raw=True, non_reg=True)
    ref = orc.calc_dcd(nump(d1), nump(d2), nump(i1), nump(i2), alpha=40.0, n_lambda=0.5, non_reg=True)
    assert_close(nump(loss), ref, what="calc_dcd non_reg")
    plain = calc_dcd(a, b, alpha=40, n_lambda=0.5)
    assert float((plain - loss).abs().max()) > 1e-4        # the clamp changes the result when n != m


@pytest.mark.parametrize("optimizer", ["ranger", "adam"])
def test_train_step_runs_and_learns(optimizer):
    """the synthetic RL_TDA step (train_step.py): finite loss, every trainable parameter that is on the path gets a
    finite gradient, and a few steps on one fixed batch reduce the loss."""
    from tgpose_b200.posenet import PoseNet9D
    from tgpose_b200.train_step import TrainStep, augment, synthetic_targets
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).cuda()
    net2 = PoseNet9D(only_encoder=True).cuda()                     # RL_TDA.py:20: frozen encoder on the augmented cloud
    step = TrainStep(net, lr=2e-4, optimizer=optimizer, net2=net2, total_iters=100)
    step.sched.kw.update(warmup_iters=0)                           # (the flags' 1000-iteration warm-up would freeze 13 steps)
    step.sched._apply()
    gen = torch.Generator().manual_seed(3)
    pts = (torch.rand(4, 256, 3, generator=gen) - 0.5) * 0.3 + torch.tensor([0.1, -0.1, 1.0])
    cat = torch.randint(0, 6, (4, 1), generator=gen).float()
    tgt = synthetic_targets(4, 5, "cuda")
    aug = augment(pts, 9).cuda()
    w2 = {n: p.detach().clone() for n, p in net2.named_parameters()}
    first = float(step(pts.cuda(), cat.cuda(), tgt, aug))
    assert set(step.last_losses) >= {"RL_loss", "recon_1_loss", "recon_consistency_loss", "R_DCD", "recon", "pose"}
    assert all(torch.isfinite(v).all() for v in step.last_losses.values())
    assert all(p.grad is None for p in net2.parameters())           # net2 runs under no_grad and is never stepped
    assert all(torch.equal(w2[n], p.detach()) for n, p in net2.named_parameters())
    for n, p in net.named_parameters():
        if "proj_layer" in n:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    hist = [first]
    for _ in range(12):
        hist.append(float(step(pts.cuda(), cat.cuda(), tgt, aug)))
    print("train-step losses:", " ".join(f"{v:.4f}" for v in hist))
    assert all(np.isfinite(v) for v in hist)
    # dropout (p = 0.5 / 0.2), 4-cloud BatchNorm statistics and fp32 atomics make the trajectory noisy: the check is that
    # optimisation makes progress at some point, not that the loss is monotone
    assert min(hist[1:]) < first, hist


@pytest.mark.parametrize("M,cin,cout,slope", [(4112, 1286, 1024, 0.0), (2056, 1024, 256, 0.0), (1000, 70, 64, 0.2)])
def test_conv_bn_act_train_vs_torch(M, cin, cout, slope):
    """heads_train.conv_bn_act_train (train-mode Conv1d(k=1) + BatchNorm1d + ReLU/LeakyReLU on the library's kernels)
    against the same torch layers in fp32 (TF32 off): output, running statistics, and every gradient."""
    from tgpose_b200.heads_train import conv_bn_act_train
    import torch.nn as nn
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + cin)
    conv_a, bn_a = nn.Conv1d(cin, cout, 1).cuda(), nn.BatchNorm1d(cout).cuda()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-0.5, 0.5)
    import copy
    conv_b, bn_b = copy.deepcopy(conv_a), copy.deepcopy(bn_a)
    x = torch.randn(M, cin).cuda()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    G = torch.randn(M, cout).cuda()
    ya = conv_bn_act_train(xa, conv_a, bn_a, slope)
    (ya * G).sum().backward()
    zb = bn_b(conv_b(xb.t().unsqueeze(0)))               # (1, cin, M) -> (1, cout, M)
    yb = (torch.relu(zb) if slope == 0.0 else torch.nn.functional.leaky_relu(zb, slope))[0].t()
    (yb * G).sum().backward()
    assert_close(nump(ya), nump(yb), rel=1e-4, floor=2e-5, what="conv+bn+act forward")
    assert_close(nump(bn_a.running_mean), nump(bn_b.running_mean), what="running_mean")
    assert_close(nump(bn_a.running_var), nump(bn_b.running_var), what="running_var")
    assert int(bn_a.num_batches_tracked) == 1
    # the activation is discontinuous in its derivative at 0: an output within rounding distance of 0 may take the other
    # branch in one of the two implementations, which changes that ROW of dx and that COLUMN of dW/dgamma/dbeta by O(1)
    # terms.  Such elements (|BN output| < 1e-5, a handful out of millions) are excluded by row / column.
    near0 = (zb[0].t().abs() < 1e-5)
    good_rows = nump(~near0.any(dim=1))
    good_cols = nump(~near0.any(dim=0))
    assert good_rows.mean() > 0.99 and good_cols.mean() > 0.9
    assert_grad_close(nump(xa.grad)[good_rows], nump(xb.grad)[good_rows], what="dx")
    assert_grad_close(nump(conv_a.weight.grad)[good_cols, :, 0], nump(conv_b.weight.grad)[good_cols, :, 0], what="dW")
    assert_grad_close(nump(bn_a.weight.grad)[good_cols], nump(bn_b.weight.grad)[good_cols], what="dgamma")
    assert_grad_close(nump(bn_a.bias.grad)[good_cols], nump(bn_b.bias.grad)[good_cols], what="dbeta")
    # the conv bias gradient is a sum of terms that cancel exactly in theory (BatchNorm removes the mean)
    assert float(conv_a.bias.grad.abs().max()) <= 1e-3 * float(G.abs().sum(0).max())
