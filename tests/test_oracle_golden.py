"""Pins oracle/ (the CPU restatement) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only; runs everywhere."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as orc
from util import assert_close, assert_knn_equal_mod_ties, golden, knn_feat_mismatch


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", list("abcdef"))
def test_knn_xyz_bit_exact(tag):
    g = golden("knn_xyz")
    x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
    idx, dist = orc.knn_xyz(x, k, return_dist=True)
    # distance matrix bit-identical to torch.bmm-based reference (gcn3d.py:18-20)
    assert sha(dist) == str(g[f"{tag}_dist_sha"])
    ref = g[f"{tag}_idx"].astype(np.int64)
    assert_knn_equal_mod_ties(idx, ref, dist, f"knn_xyz[{tag}]")
    if tag in "abde":  # unit-cube clouds: no exact distance ties -> plain equality
        assert np.array_equal(idx, ref)


def test_knn_xyz_duplicates_tie_groups():
    g = golden("knn_xyz")
    x, k = g["dup_x"], int(g["dup_k"])
    idx, dist = orc.knn_xyz(x, k, return_dist=True)
    assert np.array_equal(dist.view(np.uint32), g["dup_dist"].view(np.uint32))
    assert_knn_equal_mod_ties(idx, g["dup_idx"], dist, "knn dup")
    # lowest index first inside exact-tie groups
    d = np.take_along_axis(dist, idx, axis=2)
    same = d[..., 1:] == d[..., :-1]
    assert (idx[..., 1:][same] > idx[..., :-1][same]).all()


@pytest.mark.parametrize("tag,D", [("a", 32), ("b", 128), ("c", 256)])
def test_knn_feat(tag, D):
    g = golden("knn_feat")
    x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
    idx = orc.knn_feat(x, k)
    q = (x.astype(np.float64) ** 2).sum(-1)
    any_, set_, viol = knn_feat_mismatch(idx, g[f"{tag}_idx"], g[f"{tag}_dist"], D, q)
    assert viol == 0
    assert any_ <= 0.03 * idx.shape[0] * idx.shape[1]


@pytest.mark.parametrize("tag", list("abcd"))
def test_nearest(tag):
    g = golden("nearest")
    idx = orc.get_nearest_index(g[f"{tag}_t"], g[f"{tag}_s"])
    assert np.array_equal(idx, g[f"{tag}_idx"].astype(np.int64))


def test_gather_and_dirnorm():
    g = golden("gather_dir")
    idx = g["idx"].astype(np.int64)
    assert np.array_equal(orc.knn_xyz(g["x"], 8), idx)
    assert sha(orc.indexing_neighbor(g["f"], idx)) == str(g["gathered_sha"])
    assert_close(orc.direction_norm(g["x"], idx), g["dirs"], what="dirs")
    idx2 = g["idx2"].astype(np.int64)
    d2 = orc.direction_norm(g["x2"], idx2)
    assert_close(d2, g["dirs2"], what="dirs dup")
    assert np.isfinite(d2).all()


def _params(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def test_surface_conv():
    g = golden("convs")
    p = _params(g, "s_p_")
    x, k = g["s_x"], int(g["s_k"])
    idx = orc.knn_xyz(x, k)
    assert np.array_equal(idx, g["s_idx"].astype(np.int64))
    assert_close(orc.surface_conv(x, idx, p["directions"], 7, 16), g["s_graph"], what="surface graph_conv")
    assert_close(orc.hs_surface_forward(p, x, k), g["s_fwd"], what="surface fwd")


@pytest.mark.parametrize("tag,xk,cin,cout", [("l", "s_x", 16, 32), ("m", "m_x", 32, 16)])
def test_layer_conv(tag, xk, cin, cout):
    g = golden("convs")
    p = _params(g, f"{tag}_p_")
    x, fm = g[xk], g[f"{tag}_fm"]
    k = int(g["s_k"]) if tag == "l" else int(g["m_k"])
    idx = g[f"{tag}_idx"].astype(np.int64)
    mine = orc.knn_feat(fm, k)
    assert (mine != idx).any(axis=2).mean() < 0.03
    P = orc.gemm_bias(fm, p["weights"], p["bias"])
    if tag == "l":
        assert_close(P, g["l_proj"], what="projection")
    graph = orc.layer_conv(x, idx, p["directions"], P, 7, cout)
    assert_close(graph, g[f"{tag}_graph"], what="layer graph_conv")
    idx_orl = g[f"{tag}_idx_orl"].astype(np.int64)
    if tag == "l":
        assert_close(orc.orl_global(graph, idx_orl), g["l_orl_g"], what="ORL global")
    assert_close(orc.hs_layer_forward(p, x, fm, k, idx, idx_orl), g[f"{tag}_fwd"], what="layer fwd")


def test_pool():
    g = golden("convs")
    perm = g["p_perm"].astype(np.int64)[:32]
    v, f = orc.pool_forward(g["s_x"], g["l_fm"], perm, 4)
    assert np.array_equal(v, g["p_v"])
    assert np.array_equal(f, g["p_f"])  # pure gather + max: bit-exact


@pytest.mark.parametrize("tag", list("abc"))
def test_chamfer_forward(tag):
    """restates losses/metrics/CD/unit_test.py:14-35: dist MSE < 1e-8 and idx exactly equal."""
    g = golden("chamfer")
    for contract in (True, False):
        d1, d2, i1, i2 = orc.chamfer_forward(g[f"{tag}_p1"], g[f"{tag}_p2"], contract)
        assert np.mean((d1 - g[f"{tag}_d1"]) ** 2) + np.mean((d2 - g[f"{tag}_d2"]) ** 2) < 1e-8
        assert np.array_equal(i1, g[f"{tag}_i1"].astype(np.int32))
        assert np.array_equal(i2, g[f"{tag}_i2"].astype(np.int32))
        assert_close(d1, g[f"{tag}_d1"], what="dist1")
        assert_close(d2, g[f"{tag}_d2"], what="dist2")


def test_chamfer_backward():
    g = golden("chamfer")
    d1, d2, i1, i2 = orc.chamfer_forward(g["a_p1"], g["a_p2"])
    g1, g2 = orc.chamfer_backward(g["a_p1"], g["a_p2"], g["a_w1"], g["a_w2"], i1, i2)
    assert_close(g1, g["a_g1"], rel=1e-4, floor=1e-6, what="gradxyz1")
    assert_close(g2, g["a_g2"], rel=1e-4, floor=1e-6, what="gradxyz2")


def test_face_enc_injected_indices():
    """T2 (SURVEY 8c'): replay the reference's 14 index tensors, require feat within rel 1e-4.
    Weights: rebuilt from the seed by our own Face_Enc mirror; hashes pinned by the golden file."""
    torch = pytest.importorskip("torch")
    from tgpose_b200.face_enc import Face_Enc
    g = golden("face_enc")
    torch.manual_seed(0)
    enc = Face_Enc().eval()
    sd = {k: v.detach().numpy() for k, v in enc.state_dict().items()}
    names = [str(n) for n in g["param_names"]]
    assert sorted(sd.keys()) == names
    for n, h in zip(names, g["param_sha"]):
        assert sha(sd[n]) == str(h), f"init of {n} differs from the reference under the same seed"
    inject = [g[f"idx_{i:02d}"].astype(np.int64) for i in range(14)]
    feat = orc.face_enc_forward(sd, g["pts"], g["cat_id"], g["perm1"].astype(np.int64),
                                g["perm2"].astype(np.int64), inject=inject)
    assert_close(feat, g["feat"], what="Face_Enc feat (injected idx)")
    # free-running (T3): report only; kNN on xyz must still be exact
    feat_free = orc.face_enc_forward(sd, g["pts"], g["cat_id"], g["perm1"].astype(np.int64),
                                     g["perm2"].astype(np.int64))
    from util import frac_close
    assert frac_close(feat_free, g["feat"]) > 0.90


def test_posenet_injected_indices():
    """full network (heads included) through the oracle vs the reference's PoseNet9D outputs."""
    torch = pytest.importorskip("torch")
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).eval()
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    for n, h in zip([str(n) for n in g["param_names"]], g["param_sha"]):
        assert sha(sd[n]) == str(h), f"init of {n} differs from the reference under the same seed"
    torch.manual_seed(7)
    perm1, perm2 = torch.randperm(128).numpy(), torch.randperm(32).numpy()
    inject = [g[f"idx_{i:02d}"].astype(np.int64) for i in range(14)]
    out = orc.posenet_forward(sd, g["pts"], g["cat_id"], perm1, perm2, inject=inject)
    for k in ("recon", "p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s", "h1", "h2", "feat_global"):
        assert_close(out[k], g["out_" + k], rel=2e-4, floor=2e-6, what=k)


# ----------------------------------------------------------------------------------------- DCD loss tail
@pytest.mark.parametrize("tag", list("abcd"))
def test_calc_dcd_oracle_vs_reference(tag):
    """oracle.calc_dcd / calc_cd pinned to the reference's own calc_dcd / calc_cd (losses/TDA_loss_sym_recon.py:411-450,
    :495-509, executed from the reference source by make_golden.py: dcd_case with chamfer_python.distChamfer)."""
    g = golden("dcd")
    alpha, lam, non_reg = float(g[f"{tag}_kw"][0]), float(g[f"{tag}_kw"][1]), bool(g[f"{tag}_kw"][2])
    d1, d2, i1, i2 = orc.chamfer_forward(g[f"{tag}_pred"], g[f"{tag}_gt"])
    loss = orc.calc_dcd(d1, d2, i1, i2, alpha=alpha, n_lambda=lam, non_reg=non_reg)
    assert_close(loss, g[f"{tag}_loss"], what=f"calc_dcd {tag}")
    cd_p, cd_t = orc.calc_cd(d1, d2)
    assert_close(cd_p, g[f"{tag}_cd_p"], what="cd_p")
    assert_close(cd_t, g[f"{tag}_cd_t"], what="cd_t")


def test_posenet_1028_injected_indices():
    """the oracle at the BENCHMARKED cloud size (4 x 1028) against the reference's PoseNet9D with its indices replayed."""
    torch = pytest.importorskip("torch")
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet_1028")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).eval()
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    torch.manual_seed(7)
    perm1, perm2 = torch.randperm(1028).numpy(), torch.randperm(257).numpy()
    inject = [g[f"idx_{i:02d}"].astype(np.int64) for i in range(14)]
    orc.USE_BLAS = True
    try:
        out = orc.posenet_forward(sd, g["pts"], g["cat_id"], perm1, perm2, inject=inject)
    finally:
        orc.USE_BLAS = False
    rows = g["feat_rows"].astype(np.int64)
    assert_close(out["feat"][:, rows], g["out_feat"], what="feat rows")
    for k in ("recon", "f_green_R", "f_red_R", "Pred_T", "Pred_s", "h1", "h2", "feat_global"):
        assert_close(out[k], g["out_" + k], rel=2e-4, floor=2e-6, what=k)


# ----------------------------------------------------------------------------------------- backward oracle
def _bparams(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def test_backward_oracle_vs_reference_autograd():
    """oracle.hs_*_backward / gather_max_backward (SURVEY 8a' closed forms) against gradients produced by the
    unmodified reference's autograd (make_golden.py: backward_cases)."""
    from util import assert_grad_close
    g = golden("backward")
    x, k = g["x"], int(g["k"])
    idx_xyz = g["idx_xyz"].astype(np.int64)
    sg = orc.hs_surface_backward(_bparams(g, "s_p_"), x, k, idx_xyz, idx_xyz, g["s_G"])
    for n, v in sg.items():
        assert_grad_close(v, g["s_g_" + n], what=f"surface {n}")
    dfm, lg = orc.hs_layer_backward(_bparams(g, "l_p_"), x, g["l_fm"], k, g["l_idx"].astype(np.int64), idx_xyz, g["l_G"])
    assert_grad_close(dfm, g["l_dfm"], what="layer d_fm")
    for n, v in lg.items():
        assert_grad_close(v, g["l_g_" + n], what=f"layer {n}")
    perm = g["p_perm"].astype(np.int64)
    idx4 = orc.knn_xyz(x, 4)
    df = orc.gather_max_backward(g["p_f"], idx4, g["p_G"], rows=perm[:32])
    assert_grad_close(df, g["p_df"], what="pool d_f")


# ----------------------------------------------------------------------------------------- optimiser step
RANGER_CONFIGS = {"default": dict(lr=1e-3),
                  "wd_convonly": dict(lr=3e-3, weight_decay=0.01, gc_conv_only=True, betas=(0.9, 0.99), k=4, alpha=0.3),
                  "nogc": dict(lr=1e-2, use_gc=False, eps=1e-8),
                  "gc_update": dict(lr=2e-3, gc_loc=False, weight_decay=0.005)}


@pytest.mark.parametrize("tag", list(RANGER_CONFIGS))
def test_ranger_oracle_vs_reference(tag):
    """oracle.ranger_step against clip_grad_norm_(5) + the reference's Ranger.step() (ranger2020.py), 9 steps: both RAdam
    branches, Lookahead, weight decay incl. the reference's exp_avg alias, gc_conv_only, no centralisation."""
    z = golden("ranger")
    n, steps = int(z["n_tensors"]), int(z["steps"])
    st = orc.ranger_init([z[f"p0_{i}"] for i in range(n)])
    for t in range(steps):
        tot = orc.ranger_step(st, [z[f"g_{t}_{i}"] for i in range(n)], max_norm=5, **RANGER_CONFIGS[tag])
        assert abs(tot - z[f"{tag}_norms"][t]) <= 1e-5 * tot
        if t in z["snaps"]:
            for i in range(n):
                assert np.abs(st["p"][i] - z[f"{tag}_p_{t}_{i}"]).max() <= 2e-7, (tag, t, i)
    for name in ("m", "v", "slow"):
        for i in range(n):
            ref = z[f"{tag}_{name}_{i}"]
            assert np.abs(st[name][i] - ref).max() <= 2e-6 * (np.abs(ref).max() + 1e-30), (tag, name, i)
