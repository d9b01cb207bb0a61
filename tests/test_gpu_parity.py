"""GPU parity tests: the CUDA path (through the C-ABI, via tgpose_b200.ops / the drop-in modules)
against the CPU oracle and the golden vectors of the unmodified reference.

Bar (DESIGN.md "Parity"): xyz kNN / nearest / chamfer indices bit-exact; feature-space kNN exact up to
the rounding bound of the expanded formula; floats within rel 1e-4 (+1e-6 floor)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from util import assert_close, assert_knn_equal_mod_ties, frac_close, golden, knn_feat_mismatch

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


# ----------------------------------------------------------------------------------------- kNN
@pytest.mark.parametrize("tag", list("abcdef"))
def test_knn_xyz_golden(ops, tag):
    g = golden("knn_xyz")
    x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
    i64, i32 = ops.knn_xyz(cu(x), k, want64=True, want32=True)
    mine = nump(i64)
    assert np.array_equal(mine, nump(i32).astype(np.int64))
    ref_idx, dist = orc.knn_xyz(x, k, return_dist=True)
    assert np.array_equal(mine, ref_idx)                     # bit-exact incl. tie order (lowest index first)
    assert_knn_equal_mod_ties(mine, g[f"{tag}_idx"].astype(np.int64), dist, f"knn_xyz[{tag}] vs reference")


def test_knn_xyz_duplicates(ops):
    g = golden("knn_xyz")
    x, k = g["dup_x"], int(g["dup_k"])
    mine = nump(ops.knn_xyz(cu(x), k)[0])
    assert np.array_equal(mine, orc.knn_xyz(x, k))
    assert_knn_equal_mod_ties(mine, g["dup_idx"].astype(np.int64), g["dup_dist"], "dup vs reference")


@pytest.mark.parametrize("B,N,k", [(8, 1028, 20), (3, 257, 20), (5, 64, 8), (2, 1028, 4), (2, 333, 50), (1, 5000, 10),
                                   (2, 40, 39), (1, 2, 1), (3, 1028, 31), (2, 40, 31), (1, 33, 31), (1, 9000, 20),
                                   (1, 16500, 31), (2, 1028, 1), (2, 65, 30),
                                   # k = 32..63 on the threshold path (128 group minima)
                                   (4, 1028, 50), (2, 1028, 40), (2, 257, 63), (1, 9000, 50), (2, 64, 63), (3, 52, 50), (2, 1028, 32)])
def test_knn_xyz_vs_oracle(ops, B, N, k):
    g = torch.Generator().manual_seed(B * 1000 + N + k)
    x = torch.rand(B, N, 3, generator=g)
    mine = nump(ops.knn_xyz(x.cuda(), k)[0])
    assert np.array_equal(mine, orc.knn_xyz(x.numpy(), k))


def test_knn_xyz_half_cloud_one_point(ops):
    """PcRandomDropout (datasets/data_augmentation.py:95-97): huge tie groups, zero direction vectors."""
    g = torch.Generator().manual_seed(5)
    x = torch.rand(4, 1028, 3, generator=g)
    x[:, 514:] = x[:, :1]
    mine = nump(ops.knn_xyz(x.cuda(), 20)[0])
    assert np.array_equal(mine, orc.knn_xyz(x.numpy(), 20))
    d = nump(ops.direction_norm(x.cuda(), torch.as_tensor(mine).cuda()))
    assert np.isfinite(d).all()
    assert_close(d, orc.direction_norm(x.numpy(), mine), what="dirs with duplicates")


def test_knn_xyz_multi_tile_with_duplicates(ops):
    """N > 8192 (both selection passes walk shared-memory tiles) with tie groups larger than the survivor buffer
    (slow path inside the threshold-selection kernel), plus a cloud of identical points."""
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 9000, 3, generator=g)
    x[0, 4000:4400] = x[0, 17]
    x[1, :] = x[1, 0]
    mine = nump(ops.knn_xyz(x.cuda(), 20)[0])
    assert np.array_equal(mine, orc.knn_xyz(x.numpy(), 20))
    mine = nump(ops.knn_xyz(x.cuda(), 45)[0])            # k > 31: the two-register insertion lists of the slow path
    assert np.array_equal(mine, orc.knn_xyz(x.numpy(), 45))


def test_knn_errors(ops):
    x = torch.rand(1, 8, 3).cuda()
    with pytest.raises(RuntimeError):
        ops.knn_xyz(x, 8)          # k+1 > N: torch.topk raises in the reference
    with pytest.raises(RuntimeError):
        ops.knn_xyz(torch.rand(1, 8, 3), 2)   # CPU tensor: no fallback


@pytest.mark.parametrize("tag,D", [("a", 32), ("b", 128), ("c", 256)])
def test_knn_feat_golden(ops, tag, D):
    g = golden("knn_feat")
    x, k = g[f"{tag}_x"], int(g[f"{tag}_k"])
    mine = nump(ops.knn_feat(cu(x), k)[0])
    q = (x.astype(np.float64) ** 2).sum(-1)
    any_, set_, viol = knn_feat_mismatch(mine, g[f"{tag}_idx"], g[f"{tag}_dist"], D, q)
    assert viol == 0, f"{viol} rows differ from the reference beyond the rounding bound"
    assert any_ <= 0.03 * mine.shape[0] * mine.shape[1]


@pytest.mark.parametrize("B,N,D,k", [(4, 1028, 128, 20), (4, 257, 256, 20), (8, 64, 256, 8), (2, 130, 20, 5), (1, 4500, 64, 30),
                                     (1, 300, 64, 40), (2, 257, 128, 20), (3, 100, 128, 10), (2, 64, 64, 8), (5, 200, 128, 31),
                                     (40, 257, 64, 20),
                                     # k = 32..63: tensor-core kernel with 128 group minima (BASELINE configs[3] sweeps k to 50)
                                     (3, 1028, 128, 50), (2, 1028, 128, 40), (2, 257, 256, 63), (1, 2048, 128, 50), (2, 70, 64, 50),
                                     (2, 600, 128, 32)])
def test_knn_feat_vs_oracle(ops, B, N, D, k):
    g = torch.Generator().manual_seed(N + D)
    x = torch.randn(B, N, D, generator=g) * 0.5
    mine = nump(ops.knn_feat(x.cuda(), k)[0])
    ref, dist = orc.knn_feat(x.numpy(), k, return_dist=True)
    q = (x.numpy().astype(np.float64) ** 2).sum(-1)
    any_, set_, viol = knn_feat_mismatch(mine, ref, dist, D, q)
    assert viol == 0
    assert any_ <= 0.01 * B * N, f"{any_} of {B * N} rows differ"
    # structural: no duplicates within a row, all in range
    assert mine.min() >= 0 and mine.max() < N
    srt = np.sort(mine, axis=2)
    assert (srt[..., 1:] != srt[..., :-1]).all()


@pytest.mark.parametrize("B,N,D,k", [(3, 1028, 128, 20), (2, 300, 256, 20), (2, 64, 256, 8),
                                     (2, 1028, 128, 50), (2, 300, 64, 40)])      # k > 31: fix-up through the fp32 tile kernel
def test_knn_feat_duplicate_rows_take_the_fixup_path(ops, B, N, D, k):
    """More identical feature rows than the survivor buffer holds (zero-padded / dropped-out clouds): the threshold-selection
    kernel hands those units to the insertion-list fix-up.  Distances between duplicates are exact ties (bit-equal inner
    products), so the expected neighbours of a duplicated row are the lowest-index duplicates."""
    g = torch.Generator().manual_seed(N + D + k)
    x = torch.randn(B, N, D, generator=g) * 0.5
    ndup = min(N // 2, 150)
    x[:, N - ndup:] = x[:, :1]                                  # rows N-ndup.. equal row 0
    mine = nump(ops.knn_feat(x.cuda(), k)[0])
    ref, dist = orc.knn_feat(x.numpy(), k, return_dist=True)
    q = (x.numpy().astype(np.float64) ** 2).sum(-1)
    any_, set_, viol = knn_feat_mismatch(mine, ref, dist, D, q)
    assert viol == 0
    assert mine.min() >= 0 and mine.max() < N
    srt = np.sort(mine, axis=2)
    assert (srt[..., 1:] != srt[..., :-1]).all()
    # a duplicated row: rank 0 is dropped positionally, the k reported neighbours are the next lowest-index duplicates
    dups = np.concatenate([[0], np.arange(N - ndup, N)])
    for b in range(B):
        got = mine[b, N - 1]
        assert set(got.tolist()) <= set(dups.tolist()), got
        assert (np.diff(got) > 0).all()                         # ties in ascending index order


@pytest.mark.parametrize("tag", list("abcd"))
def test_nearest_golden(ops, tag):
    g = golden("nearest")
    mine = nump(ops.nearest(cu(g[f"{tag}_t"]), cu(g[f"{tag}_s"]))[0])
    assert np.array_equal(mine, g[f"{tag}_idx"].astype(np.int64))


def test_nearest_vs_oracle_large(ops):
    g = torch.Generator().manual_seed(3)
    t, s = torch.rand(6, 1028, 3, generator=g), torch.rand(6, 2500, 3, generator=g)
    assert np.array_equal(nump(ops.nearest(t.cuda(), s.cuda())[0]), orc.get_nearest_index(t.numpy(), s.numpy()))


# ----------------------------------------------------------------------------------------- gathers
def test_gather_dir_select(ops):
    g = golden("gather_dir")
    idx = g["idx"].astype(np.int64)
    out = nump(ops.gather_rows(cu(g["f"]), cu(idx)))
    assert np.array_equal(out, orc.indexing_neighbor(g["f"], idx))
    out32 = nump(ops.gather_rows(cu(g["f"]), cu(idx.astype(np.int32))))
    assert np.array_equal(out, out32)
    f5 = np.random.default_rng(0).standard_normal((2, 128, 5)).astype(np.float32)   # C % 4 != 0 path
    assert np.array_equal(nump(ops.gather_rows(cu(f5), cu(idx))), orc.indexing_neighbor(f5, idx))
    assert_close(nump(ops.direction_norm(cu(g["x"]), cu(idx))), g["dirs"], what="dirs")
    assert_close(nump(ops.direction_norm(cu(g["x2"]), cu(g["idx2"].astype(np.int64)))), g["dirs2"], what="dirs dup")
    rows = np.array([5, 0, 127, 64, 5], np.int64)
    assert np.array_equal(nump(ops.select_rows(cu(g["f"]), cu(rows))), g["f"][:, rows, :])


@pytest.mark.parametrize("C", [128, 24, 7])
def test_gather_max_and_orl(ops, C):
    rng = np.random.default_rng(C)
    B, N, k = 3, 257, 20
    x = rng.random((B, N, 3), dtype=np.float32)
    f = rng.standard_normal((B, N, C)).astype(np.float32)
    idx = orc.knn_xyz(x, k)
    out, arg = ops.gather_max(cu(f), cu(idx), want_arg=True)
    ref, ref_arg = orc.gather_max(f, idx, return_arg=True)
    assert np.array_equal(nump(out), ref)
    assert np.array_equal(nump(arg), ref_arg)
    rows = np.sort(rng.permutation(N)[:64]).astype(np.int64)
    assert np.array_equal(nump(ops.gather_max(cu(f), cu(idx.astype(np.int32)), rows=cu(rows))), orc.gather_max(f, idx, rows))
    gl, garg = ops.orl_global(cu(f), cu(idx.astype(np.int32)), want_arg=True)
    assert_close(nump(gl), orc.orl_global(f, idx), what="orl_global")
    assert np.array_equal(nump(garg), ref_arg)


# ----------------------------------------------------------------------------------------- gemm
@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("M,K,N,nk", [(300, 128, 1024, False), (257, 3, 128, True), (1000, 256, 256, True),
                                      (64, 128, 130, False), (5, 16, 7, True), (129, 20, 36, False),
                                      (4112, 1289, 520, True), (700, 100, 72, False), (32, 1286, 512, True),
                                      (3, 256, 40, True), (2000, 128, 3, True), (2000, 3, 128, False),
                                      (32, 5000, 1286, True), (40, 1024, 5000, True), (64, 200, 33, True),
                                      (17, 129, 19, True), (1, 33, 1, True)])
def test_gemm_plain(ops, M, K, N, nk, tc):
    """both contraction kernels (fp32 FMA and tcgen05 3xTF32) against the fp64-accumulated oracle.
    Tolerance: rel 1e-4 with an absolute floor of 1e-5 (sums of K unit-scale products)."""
    if tc and K > 4096:
        pytest.skip("K = 5000 only occurs in per-cloud (M = batch) contractions, which are pinned to the fp32 kernel")
    rng = np.random.default_rng(M + K + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32) * 0.1
    bias = rng.standard_normal(N).astype(np.float32)
    out = torch.empty(M, N, device="cuda")
    Bm = cu(np.ascontiguousarray(W.T)) if nk else cu(W)
    ops.gemm(cu(A), Bm, nk, [(0, N, out, 0, 0)], bias=cu(bias), tc=tc)
    assert_close(nump(out), orc.gemm_bias(A, W, bias), rel=1e-4, floor=1e-5 * max(1.0, (K / 128) ** 0.5), what="gemm")


@pytest.mark.parametrize("tc,Np", [(False, 70), (True, 70), (True, 300)])
def test_gemm_epilogue_and_segments(ops, tc, Np):
    rng = np.random.default_rng(11)
    Bc, K, C, S = 3, 32, 16, 7
    M = Bc * Np
    A = rng.standard_normal((M, K)).astype(np.float32)
    Ncols = (S + 2) * C
    W = rng.standard_normal((K, Ncols)).astype(np.float32) * 0.2
    bias = rng.standard_normal(Ncols).astype(np.float32)
    gb = rng.standard_normal((Bc, Ncols)).astype(np.float32)
    r1 = rng.standard_normal((M, Ncols)).astype(np.float32)
    r2 = rng.standard_normal((M, Ncols)).astype(np.float32)
    sc = rng.random(Ncols).astype(np.float32) + 0.5
    sh = rng.standard_normal(Ncols).astype(np.float32)
    centre = torch.zeros(M, C, device="cuda")
    slab = torch.zeros(C // 4, M, S * 4, device="cuda")
    ste = torch.zeros(M, C + 3, device="cuda")[:, :C]      # row-strided destination
    Kp_out = 32
    spl = torch.zeros(M, 2 * Kp_out, device="cuda")          # split copy of the centre columns (mode 2)
    ops.gemm(cu(A), cu(W), False, [(0, C, centre, 0, 0), (C, C + S * C, slab, 1, S * 4), (C + S * C, Ncols, ste, 0, 0),
                                   (0, C, spl, 2, Kp_out)],
             bias=cu(bias), group_bias=cu(gb), rows_per_group=Np, res1=cu(r1), res2=cu(r2), scale=cu(sc), shift=cu(sh),
             relu=True, tc=tc)
    ref = orc.gemm_bias(A, W, bias).astype(np.float64) + np.repeat(gb, Np, axis=0) + r1 + r2
    ref = np.maximum(ref * sc + sh, 0).astype(np.float32)
    assert_close(nump(centre), ref[:, :C], rel=1e-4, floor=1e-5, what="centre")
    assert_close(nump(ste), ref[:, C + S * C:], rel=1e-4, floor=1e-5, what="ste")
    sl = ref[:, C:C + S * C].reshape(M, C // 4, S * 4).transpose(1, 0, 2)
    assert_close(nump(slab), sl, rel=1e-4, floor=1e-5, what="slab")
    sp = nump(spl)
    assert np.array_equal(sp[:, :C] + sp[:, Kp_out:Kp_out + C], nump(centre))          # hi + lo == value, exactly
    assert (sp[:, :C].view(np.uint32) & 0x1FFF == 0).all()                              # hi is a tf32 number
    assert (sp[:, C:Kp_out] == 0).all() and (sp[:, Kp_out + C:] == 0).all()
    # leaky slope per column
    out2 = torch.empty(M, Ncols, device="cuda")
    slope = rng.random(Ncols).astype(np.float32)
    ops.gemm(cu(A), cu(W), False, [(0, Ncols, out2, 0, 0)], bias=cu(bias), neg_slope=cu(slope), tc=tc)
    r2_ = orc.gemm_bias(A, W, bias)
    assert_close(nump(out2), np.where(r2_ > 0, r2_, r2_ * slope), rel=1e-4, floor=1e-5, what="leaky")


# ----------------------------------------------------------------------------------------- convs
def _params(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def _slab_from_P(P, S, C):
    """(B,N,(S+1)C) projected features -> (centre (M,C), slab [C/4][M][S*4]) as the GEMM would write them."""
    M = P.shape[0] * P.shape[1]
    P2 = P.reshape(M, (S + 1) * C)
    sup = P2[:, C:].reshape(M, S, C // 4, 4).transpose(2, 0, 1, 3).reshape(C // 4, M, S * 4)
    return np.ascontiguousarray(P2[:, :C]), np.ascontiguousarray(sup)


def test_surface_conv_golden(ops):
    g = golden("convs")
    p = _params(g, "s_p_")
    x, k = g["s_x"], int(g["s_k"])
    idx = g["s_idx"].astype(np.int64)
    out, arg = ops.surface_conv(cu(x), cu(idx), cu(p["directions"]), 7, 16, want_arg=True)
    assert_close(nump(out), g["s_graph"], what="surface graph_conv vs reference")
    assert_close(nump(ops.surface_conv(cu(x), cu(idx.astype(np.int32)), cu(p["directions"]), 7, 16)), g["s_graph"],
                 what="surface (int32 idx)")
    assert nump(arg).max() < k


@pytest.mark.parametrize("S,C,N,k", [(7, 128, 1028, 20), (7, 32, 100, 9), (3, 20, 64, 5), (8, 64, 257, 40)])
def test_surface_conv_vs_oracle(ops, S, C, N, k):
    rng = np.random.default_rng(S * C)
    x = rng.random((2, N, 3), dtype=np.float32)
    dirs = (rng.random((3, S * C), dtype=np.float32) - 0.5)
    idx = orc.knn_xyz(x, k)
    assert_close(nump(ops.surface_conv(cu(x), cu(idx), cu(dirs), S, C)), orc.surface_conv(x, idx, dirs, S, C),
                 what="surface vs oracle")


@pytest.mark.parametrize("tag,xk,cout", [("l", "s_x", 32), ("m", "m_x", 16)])
def test_layer_conv_golden(ops, tag, xk, cout):
    g = golden("convs")
    p = _params(g, f"{tag}_p_")
    x, fm = g[xk], g[f"{tag}_fm"]
    idx = g[f"{tag}_idx"].astype(np.int64)
    B, N, _ = x.shape
    P = orc.gemm_bias(fm, p["weights"], p["bias"])
    centre, slab = _slab_from_P(P, 7, cout)
    rec = ops.edge_records(cu(x), cu(idx))
    out = ops.layer_conv(rec, cu(p["directions"]), cu(centre), cu(slab), B, N, 7, cout)
    assert_close(nump(out), g[f"{tag}_graph"], what="layer graph_conv vs reference")


@pytest.mark.parametrize("S,C,N,k,B", [(7, 128, 1028, 20, 2), (7, 256, 257, 20, 3), (7, 512, 64, 8, 2), (4, 8, 50, 33, 1),
                                       (7, 16, 2500, 12, 1)])     # last: N*S*16 B > shared memory -> L2-gather variant
def test_layer_conv_vs_oracle(ops, S, C, N, k, B):
    rng = np.random.default_rng(C + N)
    x = rng.random((B, N, 3), dtype=np.float32)
    dirs = rng.random((3, S * C), dtype=np.float32) - 0.5
    P = rng.standard_normal((B, N, (S + 1) * C)).astype(np.float32)
    idx = np.stack([np.stack([rng.permutation(N)[:k] for _ in range(N)]) for _ in range(B)]).astype(np.int64)
    centre, slab = _slab_from_P(P, S, C)
    rec = ops.edge_records(cu(x), cu(idx.astype(np.int32)))
    out, arg = ops.layer_conv(rec, cu(dirs), cu(centre), cu(slab), B, N, S, C, want_arg=True)
    assert_close(nump(out), orc.layer_conv(x, idx, dirs, P, S, C), what="layer vs oracle")
    out2 = ops.layer_conv(rec, cu(dirs), cu(centre), cu(slab), B, N, S, C)
    assert np.array_equal(nump(out), nump(out2))
    assert nump(arg).max() < k


# ----------------------------------------------------------------------------------------- modules
def _load(mod, p):
    sd = {k: torch.as_tensor(v) for k, v in p.items()}
    mod.load_state_dict(sd)
    return mod.cuda().eval()


def test_hs_surface_module_golden():
    from tgpose_b200 import gcn3d
    g = golden("convs")
    m = _load(gcn3d.HSlayer_surface(16, 7), _params(g, "s_p_"))
    with torch.no_grad():
        out = m(cu(g["s_x"]), int(g["s_k"]))
    assert_close(nump(out), g["s_fwd"], what="HSlayer_surface.forward vs reference")


@pytest.mark.parametrize("tag,xk,cin,cout", [("l", "s_x", 16, 32), ("m", "m_x", 32, 16)])
def test_hs_layer_module_golden(tag, xk, cin, cout):
    from tgpose_b200 import gcn3d
    g = golden("convs")
    m = _load(gcn3d.HS_layer(cin, cout, 7), _params(g, f"{tag}_p_"))
    k = int(g["s_k"]) if tag == "l" else int(g["m_k"])
    with torch.no_grad():
        # T1: reference indices injected
        out = m(cu(g[xk]), cu(g[f"{tag}_fm"]), k, idx_feat=cu(g[f"{tag}_idx"], torch.int32))
        free = m(cu(g[xk]), cu(g[f"{tag}_fm"]), k)
    assert_close(nump(out), g[f"{tag}_fwd"], what="HS_layer.forward vs reference (reference idx)")
    assert frac_close(nump(free), g[f"{tag}_fwd"]) > 0.97


def test_pool_module_golden():
    from tgpose_b200 import gcn3d
    g = golden("convs")
    pool = gcn3d.Pool_layer(4, 4)
    torch.manual_seed(7)
    with torch.no_grad():
        v, f = pool(cu(g["s_x"]), cu(g["l_fm"]))
    assert np.array_equal(nump(v), g["p_v"])
    assert np.array_equal(nump(f), g["p_f"])


def test_module_functions_match_reference_api():
    from tgpose_b200 import gcn3d
    g = golden("gather_dir")
    x = cu(g["x"])
    idx = gcn3d.get_neighbor_index(x, 8)
    assert idx.dtype == torch.int64 and np.array_equal(nump(idx), g["idx"].astype(np.int64))
    rf, idx2 = gcn3d.get_receptive_fields(8, x, mode='RF-P')
    assert_close(nump(rf), g["dirs"], what="RF-P dirs")
    f = cu(g["f"])
    assert np.array_equal(nump(gcn3d.indexing_neighbor_new(f, idx)), orc.indexing_neighbor(g["f"], g["idx"].astype(np.int64)))
    gl = gcn3d.get_ORL_global(f, x, 8)
    assert gl.shape == (2, 128, 24)
    assert_close(nump(gl[:, 0]), orc.orl_global(g["f"], g["idx"].astype(np.int64)), what="get_ORL_global")
    nn_ = gcn3d.get_nearest_index(x, x[:, :32].contiguous())
    assert nn_.shape == (2, 128, 1) and nn_.dtype == torch.int64


def test_face_enc_injected_and_free():
    """T2: the reference's 14 index tensors replayed -> feat within rel 1e-4 everywhere.
    T3: free-running -> report the in-tolerance fraction (gate only on gross failure)."""
    from tgpose_b200.face_enc import Face_Enc
    g = golden("face_enc")
    torch.manual_seed(0)
    enc = Face_Enc().cuda().eval()
    pts, cat = cu(g["pts"]), cu(g["cat_id"])
    with torch.no_grad():
        enc._inject = [cu(g[f"idx_{i:02d}"].astype(np.int32)) for i in range(14)]
        torch.manual_seed(7)
        feat, _ = enc(pts, cat)
        enc._inject = None
        enc._record = []
        torch.manual_seed(7)
        feat_free, fg = enc(pts, cat)
    assert feat.shape == (2, 128, 1286) and fg.shape == (2, 1286, 128)
    assert_close(nump(feat), g["feat"], what="Face_Enc feat with reference indices (T2)")
    # xyz-space index tensors of the free run are bit-exact vs the reference -- all ten of them (the pooled levels see
    # the same points because Pool_layer's permutation comes from the same CPU seed)
    for slot in (0, 1, 3, 4, 6, 8, 9, 11, 12, 13):
        assert np.array_equal(nump(enc._record[slot]).astype(np.int64), g[f"idx_{slot:02d}"].astype(np.int64)), slot
    frac = frac_close(nump(feat_free), g["feat"])
    print(f"T3 free-running in-tolerance fraction: {frac:.4f}")
    assert frac > 0.90


# ----------------------------------------------------------------------------------------- chamfer
@pytest.mark.parametrize("tag", list("abc"))
def test_chamfer_golden(tag):
    """losses/metrics/CD/unit_test.py:14-35 restated: MSE < 1e-8 and idx exactly equal."""
    from tgpose_b200.dist_chamfer_3D import chamfer_3DDist
    g = golden("chamfer")
    d1, d2, i1, i2 = chamfer_3DDist()(cu(g[f"{tag}_p1"]), cu(g[f"{tag}_p2"]))
    assert i1.dtype == torch.int32 and d1.dtype == torch.float32
    assert np.mean((nump(d1) - g[f"{tag}_d1"]) ** 2) + np.mean((nump(d2) - g[f"{tag}_d2"]) ** 2) < 1e-8
    assert np.array_equal(nump(i1), g[f"{tag}_i1"].astype(np.int32))
    assert np.array_equal(nump(i2), g[f"{tag}_i2"].astype(np.int32))
    o1, o2, oi1, oi2 = orc.chamfer_forward(g[f"{tag}_p1"], g[f"{tag}_p2"], contract=True)
    assert np.array_equal(nump(d1), o1) and np.array_equal(nump(d2), o2)      # bit-exact vs the fma restatement
    assert np.array_equal(nump(i1), oi1) and np.array_equal(nump(i2), oi2)


def test_chamfer_extension_api_and_backward():
    """chamfer_3D.forward/backward on caller-allocated tensors (chamfer_cuda.cpp:17-33), returns 1."""
    from tgpose_b200 import chamfer_3D
    g = golden("chamfer")
    p1, p2 = cu(g["a_p1"]), cu(g["a_p2"])
    d1, d2 = torch.zeros(4, 100, device="cuda"), torch.zeros(4, 200, device="cuda")
    i1 = torch.zeros(4, 100, device="cuda", dtype=torch.int32)
    i2 = torch.zeros(4, 200, device="cuda", dtype=torch.int32)
    assert chamfer_3D.forward(p1, p2, d1, d2, i1, i2) == 1
    g1, g2 = torch.zeros_like(p1), torch.zeros_like(p2)
    assert chamfer_3D.backward(p1, p2, g1, g2, cu(g["a_w1"]), cu(g["a_w2"]), i1, i2) == 1
    assert_close(nump(g1), g["a_g1"], what="gradxyz1 vs autograd of the pure-torch chamfer")
    assert_close(nump(g2), g["a_g2"], what="gradxyz2")
    o1, o2 = orc.chamfer_backward(g["a_p1"], g["a_p2"], g["a_w1"], g["a_w2"], nump(i1), nump(i2))
    assert_close(nump(g1), o1, what="gradxyz1 vs oracle")
    assert_close(nump(g2), o2, what="gradxyz2 vs oracle")


def test_chamfer_autograd_and_full_size_properties():
    from tgpose_b200.dist_chamfer_3D import calc_cd, chamfer_3DDist
    g = torch.Generator().manual_seed(9)
    B, n, m = 32, 1028, 1024
    a = torch.rand(B, n, 3, generator=g).cuda().requires_grad_(True)
    b = torch.rand(B, m, 3, generator=g).cuda()
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    # property: dist is exactly the squared distance to the reported point; nothing is closer
    pick = torch.gather(b, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    diff = pick - a.detach()
    assert torch.allclose(d1, (diff * diff).sum(-1), rtol=1e-5, atol=1e-9)
    full = torch.cdist(a.detach().double(), b.double()) ** 2
    assert torch.allclose(d1.double(), full.min(2)[0], rtol=1e-4, atol=1e-9)
    assert torch.allclose(d2.double(), full.min(1)[0], rtol=1e-4, atol=1e-9)
    # identical clouds: zero distance, identity index
    e1, e2, j1, j2 = chamfer_3DDist()(b, b)
    assert float(e1.abs().max()) == 0.0 and torch.equal(j1.long(), torch.arange(m, device="cuda").expand(B, -1))
    cd_p, cd_t = calc_cd(a, b)
    (cd_t.sum() + d1.sum()).backward()
    assert a.grad is not None and torch.isfinite(a.grad).all()
    o1, o2, oi1, oi2 = orc.chamfer_forward(nump(a), nump(b))
    assert np.array_equal(nump(i1), oi1) and np.array_equal(nump(i2), oi2)


# ----------------------------------------------------------------------------------------- full size
def test_full_size_encoder_properties():
    """B=32 x 1028 (BASELINE config[1] size): shapes, finiteness, batch-independence (the path shards by cloud)."""
    from tgpose_b200.face_enc import Face_Enc
    torch.manual_seed(0)
    enc = Face_Enc().cuda().eval()
    g = torch.Generator().manual_seed(1234)
    pts = torch.rand(32, 1028, 3, generator=g)
    pts = (pts - pts.mean(1, keepdim=True)).cuda()
    cat = torch.randint(0, 6, (32, 1), generator=g).float().cuda()
    with torch.no_grad():
        torch.manual_seed(7)
        feat, _ = enc(pts, cat)
        torch.manual_seed(7)
        feat_half, _ = enc(pts[8:16].contiguous(), cat[8:16].contiguous())
    assert feat.shape == (32, 1028, 1286) and torch.isfinite(feat).all()
    # a cloud's features do not depend on which batch it sits in (eval mode, same permutation seed):
    # every kernel of the path is deterministic and per-cloud, so this is bit-exact
    assert torch.equal(feat[8:16], feat_half)


def test_posenet_golden():
    """full network (heads through the GEMM epilogue path) vs the reference's PoseNet9D, indices replayed."""
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).cuda().eval()
    net.face_all.encoder._inject = [cu(g[f"idx_{i:02d}"].astype(np.int32)) for i in range(14)]
    with torch.no_grad():
        torch.manual_seed(7)
        out = net(cu(g["pts"]), cu(g["cat_id"]))
    for k in ("recon", "f_green_R", "f_red_R", "h1", "h2", "feat_global"):
        assert_close(nump(out[k]), g["out_" + k], rel=2e-4, floor=2e-6, what=k)
    # Pose vectors at rel 1e-4 (north_star): unit axes within 1e-4 rad = 0.0057 deg, T / s within 1e-5 absolute
    # (measured on B200, scripts/pose_err.py: 1.1e-4 / 7.0e-4 deg, 1.3e-7 / 4.4e-7)
    for k in ("p_green_R", "p_red_R"):
        a, b = nump(out[k]).astype(np.float64), g["out_" + k].astype(np.float64)
        cosang = np.clip((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1)), -1, 1)
        assert np.degrees(np.arccos(cosang)).max() < 0.0057, k
    for k in ("Pred_T", "Pred_s"):
        assert np.abs(nump(out[k]) - g["out_" + k]).max() < 1e-5, k


def _axis_angle_deg(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    cosang = np.clip((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1)), -1, 1)
    return float(np.degrees(np.arccos(cosang)).max())


def test_posenet_1028_golden_tensor_core_path():
    """T2 at the BENCHMARKED cloud size: 4 x 1028 points, so every level has >= 256 rows and every contraction of the
    encoder runs on tcgen05 (3xTF32), the heads on mixed operands -- against the reference's own PoseNet9D output with its
    14 index tensors replayed (tests/golden/posenet_1028.npz).  T3: the free-running fraction is printed."""
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet_1028")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).cuda().eval()
    enc = net.face_all.encoder
    pts, cat = cu(g["pts"]), cu(g["cat_id"])
    rows = g["feat_rows"].astype(np.int64)
    with torch.no_grad():
        enc._inject = [cu(g[f"idx_{i:02d}"].astype(np.int32)) for i in range(14)]
        torch.manual_seed(7)
        out = {k: v.clone() for k, v in net(pts, cat).items()}
        enc._inject = None
        enc._record = []
        torch.manual_seed(7)
        free = {k: v.clone() for k, v in net(pts, cat).items()}
        rec = [nump(t).astype(np.int64) for t in enc._record]
        enc._record = None
    assert_close(nump(out["feat"])[:, rows], g["out_feat"], what="feat (T2, 4 x 1028, tcgen05 path)")
    for k in ("recon", "f_green_R", "f_red_R", "h1", "h2", "feat_global"):
        assert_close(nump(out[k]), g["out_" + k], rel=2e-4, floor=2e-6, what=k)
    ang = {k: _axis_angle_deg(nump(out[k]), g["out_" + k]) for k in ("p_green_R", "p_red_R")}
    dts = {k: float(np.abs(nump(out[k]) - g["out_" + k]).max()) for k in ("Pred_T", "Pred_s")}
    print(f"T2 4x1028 pose error vs reference: axes {ang} deg, T/s abs {dts}")
    # poses at rel 1e-4 (north_star): unit axes within 1e-4 rad = 0.0057 deg, T / s (|.| ~ 0.1 .. 1) within 1e-5 absolute
    # (measured on B200: 1.2e-4 / 1.2e-3 deg, 2.4e-7 / 4.2e-7)
    assert max(ang.values()) < 0.0057 and max(dts.values()) < 1e-5
    # every xyz-space index tensor of the free run is bit-exact vs the reference (call order: SURVEY 8a a1):
    # 0 conv_0 RF-P, 1 conv_0 ORL, 3 conv_1 ORL, 4 pool_1, 6 conv_2 ORL, 8 conv_3 ORL, 9 pool_2, 11 conv_4 ORL, 12/13 nearest
    for slot in (0, 1, 3, 4, 6, 8, 9, 11, 12, 13):
        assert np.array_equal(rec[slot], g[f"idx_{slot:02d}"].astype(np.int64)), f"xyz index slot {slot}"
    # feature-space slots (2, 5, 7, 10): rows whose neighbour SET differs from the reference's (ill-conditioned, SURVEY 8c')
    for slot in (2, 5, 7, 10):
        a, b = np.sort(rec[slot], axis=-1), np.sort(g[f"idx_{slot:02d}"].astype(np.int64), axis=-1)
        print(f"T3 slot {slot}: {float((a != b).any(-1).mean()):.4f} of rows with a different neighbour set")
    frac = frac_close(nump(free["feat"])[:, rows], g["out_feat"])
    print(f"T3 free-running in-tolerance fraction at 4 x 1028: {frac:.4f}")
    # T3 is reported, not gated on (SURVEY 8c'): feature-space kNN is ill-conditioned, a flipped neighbour moves a
    # max-over-k, the ORL cloud mean carries it to every row and seeds more flips downstream (measured here: 0.9 % /
    # 3.9 % / 6.7 % / 10.2 % of rows with a different neighbour SET at the four RF-F calls, 62 % of feat elements inside
    # rel 1e-4).  The gate is the other direction of T2: OUR free-running indices replayed into the CPU oracle must
    # reproduce OUR outputs within rel 1e-4 -- i.e. every difference from the reference is an index flip inside the
    # rounding bound of the reference's own formula (rule 3, tested at op level), never arithmetic.
    assert frac > 0.5
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    torch.manual_seed(7)
    perm1, perm2 = torch.randperm(1028).numpy(), torch.randperm(257).numpy()
    orc.USE_BLAS = True
    try:
        ref = orc.posenet_forward(sd, g["pts"], g["cat_id"], perm1, perm2, inject=rec)
    finally:
        orc.USE_BLAS = False
    assert_close(nump(free["feat"]), ref["feat"], what="free-running feat vs oracle with OUR indices replayed")
    for k in ("recon", "h1", "h2", "feat_global", "Pred_T", "Pred_s"):
        assert_close(nump(free[k]), ref[k], rel=2e-4, floor=2e-6, what=f"free-running {k} vs oracle replay")


def test_cuda_graph_replay_equals_eager():
    """graph.GraphedPoseNet: the replayed forward (static buffers, Pool draws made on the host in the reference's
    order) is bit-identical to the eager forward for the same CPU seed, for two different inputs."""
    from tgpose_b200.graph import GraphedPoseNet
    from tgpose_b200.posenet import PoseNet9D
    torch.manual_seed(0)
    net = PoseNet9D().cuda().eval()
    B, N = 4, 1028
    gr = GraphedPoseNet(net, B, N)
    gen = torch.Generator().manual_seed(21)
    for trial in range(2):
        pts = (torch.rand(B, N, 3, generator=gen) - 0.5) * 0.3 + torch.tensor([0.0, 0.1, 1.0])
        cat = torch.randint(0, 6, (B, 1), generator=gen).float()
        torch.manual_seed(100 + trial)
        with torch.no_grad():
            eager = {k: v.clone() for k, v in net(pts.cuda(), cat.cuda()).items()}
        torch.manual_seed(100 + trial)
        out = gr(pts.pin_memory(), cat.pin_memory())
        torch.cuda.synchronize()
        for k in eager:
            assert torch.equal(eager[k], out[k]), (trial, k)


@pytest.mark.parametrize("M,K,N", [(1028, 1289, 512), (300, 64, 128), (4112, 1024, 256), (257, 200, 96)])
def test_gemm_mixed_operands(ops, M, K, N):
    """tgp_gemm with MIXED operands (fp16 hi.hi + bf16 cross terms, the heads' contraction): error vs an fp64 product
    stays at fp32-summation-noise level, and the mode-4 epilogue writes an operand that a second contraction accepts."""
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ W.double().t() + bias.double()
    out = torch.empty(M, N, device="cuda")
    mix = ops.mixed_buf(M, N, "cuda")
    ops.gemm(None, W, True, [(0, N, out, 0, 0), (0, N, mix, 4, ops.mixed_kpad(N))], bias=bias, K=K,
             A_split=ops.split_mixed(A), B_split=ops.split_mixed(W), mixed=True)
    scale = float(ref.abs().max())
    err = float((out.double() - ref).abs().max())
    # measured ~1e-6 * scale; the bound below is 2^-16 of the output scale (the fp32 FMA kernel sits at ~2^-20)
    assert err <= 1.6e-5 * scale, (err, scale)
    # the mode-4 operand equals the split of the raw output (bit for bit), so the next layer can consume it
    kp3 = 3 * ops.mixed_kpad(N)            # 16-bit slots in use: [fp16(x) | bf16(x) | bf16(x - fp16(x))]; the last quarter is unused
    assert torch.equal(mix.view(torch.int16)[:, :kp3], ops.split_mixed(out).view(torch.int16)[:, :kp3])
    W2 = (torch.randn(64, N, generator=g) * 0.1).cuda()
    out2 = torch.empty(M, 64, device="cuda")
    ops.gemm(None, W2, True, [(0, 64, out2, 0, 0)], K=N, A_split=mix, B_split=ops.split_mixed(W2), mixed=True)
    ref2 = out.double() @ W2.double().t()
    assert float((out2.double() - ref2).abs().max()) <= 1.6e-5 * float(ref2.abs().max())


@pytest.mark.parametrize("M,K,N,npg,ncoarse", [(4112, 265, 512, 1028, 257), (1028, 128, 256, 257, 64), (300, 64, 128, 100, 7),
                                                (2056, 320, 384, 1028, 257), (8224, 265, 1024, 1028, 257),
                                                (8224, 64, 576, 1028, 257)])
def test_gemm_gathered_residuals(ops, M, K, N, npg, ncoarse):
    """tgp_gemm_args.res1_idx / res2_idx: output row m adds row idx[m] of the residual matrices before the affine / activation.
    Checked against an fp64 product for every destination kind (raw, split, mixed, per-cloud column max) together with a
    per-cloud bias, on shapes with full 32-row blocks (fast chunk) and ragged last blocks, for mixed and 3xTF32 operands; the
    last two shapes run 256-column tiles, i.e. the residual pieces prefetched into shared memory (gemm_tc.cu, vec_ok == 5) --
    the very last one with 144-column segments, so that fast and general chunks alternate inside a tile and the last column
    tile is ragged."""
    g = torch.Generator().manual_seed(M + K + N)
    B = M // npg
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    R1 = torch.randn(B * ncoarse, N + 8, generator=g).cuda()[:, 4:4 + N]      # row-strided views (ld = N + 8)
    R2 = torch.randn(B * 5, N, generator=g).cuda()
    i1 = (torch.randint(0, ncoarse, (B, npg), generator=g) + torch.arange(B).view(B, 1) * ncoarse).int().cuda().reshape(-1)
    i2 = (torch.randint(0, 5, (B, npg), generator=g) + torch.arange(B).view(B, 1) * 5).int().cuda().reshape(-1)
    gb = torch.randn(B, N, generator=g).cuda()
    scale, shift = (torch.rand(N, generator=g) + 0.5).cuda(), torch.randn(N, generator=g).cuda()
    slope = torch.full((N,), 0.2).cuda()
    pre = A.double() @ W.double().t() + R1.double()[i1.long()] + R2.double()[i2.long()] + gb.double().repeat_interleave(npg, 0)
    v = pre * scale.double() + shift.double()
    ref = torch.where(v > 0, v, 0.2 * v)
    tol = 1.6e-5 * float(pre.abs().max()) * float(scale.max())
    h = N // 4
    for mixed in (True, False):
        a_split = ops.split_mixed(A) if mixed else ops.split_tf32(A)
        w_split = ops.split_mixed(W) if mixed else ops.split_tf32(W)
        raw = torch.empty(M, h, device="cuda")
        spl = ops._split_buf(M, h, "cuda")
        mix = ops.mixed_buf(M, h, "cuda")
        mx = torch.full((B, h), -2 ** 31, dtype=torch.int32, device="cuda")
        segs = [(0, h, raw, 0, 0), (h, 2 * h, spl, 2, ops.kpad(h)), (2 * h, 3 * h, mix, 4, ops.mixed_kpad(h)), (3 * h, N, mx, 3, 0)]
        ops.gemm(None, W, True, segs, K=K, A_split=a_split, B_split=w_split, mixed=mixed, group_bias=gb, rows_per_group=npg,
                 res1=R1, res2=R2, res1_idx=i1, res2_idx=i2, scale=scale, shift=shift, neg_slope=slope)
        assert float((raw.double() - ref[:, :h]).abs().max()) <= tol
        got_spl = spl[:, :h].double() + spl[:, ops.kpad(h):ops.kpad(h) + h].double()
        assert float((got_spl - ref[:, h:2 * h]).abs().max()) <= tol
        # the mixed destination: fp16 slot + residual slot reproduce the value to ~2^-19
        m16 = mix.view(torch.int16)
        kp = ops.mixed_kpad(h)
        hi = m16[:, :h].view(torch.float16).double()
        lo = m16[:, 2 * kp:2 * kp + h].view(torch.bfloat16).double()
        assert float((hi + lo - ref[:, 2 * h:3 * h]).abs().max()) <= tol + 4e-6 * float(ref.abs().max())
        got_mx = ops.decode_max(mx).double()
        assert float((got_mx - ref[:, 3 * h:].view(B, npg, -1).amax(1)).abs().max()) <= tol
    # the exact-fp32 kernel takes the same arguments
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, W, True, [(0, N, out, 0, 0)], group_bias=gb, rows_per_group=npg, res1=R1, res2=R2, res1_idx=i1, res2_idx=i2,
             scale=scale, shift=shift, neg_slope=slope, tc=False)
    assert float((out.double() - ref).abs().max()) <= tol


def test_posenet_factored_heads_equal_unfactored():
    """The heads' first layers with the upsampled channels contracted per coarse point (W.[x | up(y)] = W_x.x + up(W_y.y),
    posenet.py) against the same layers over the materialised concatenation: same outputs to fp32 summation noise."""
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet_1028")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).cuda().eval()
    pts, cat = cu(g["pts"]), cu(g["cat_id"])
    outs = {}
    with torch.no_grad():
        for fac in (True, False):
            net.factored_heads = fac
            net.face_all.encoder._record = []
            torch.manual_seed(7)
            outs[fac] = {k: v.clone() for k, v in net(pts, cat).items()}
            if fac:
                net.face_all.encoder._inject = [t.clone() for t in net.face_all.encoder._record]   # same neighbourhoods in both runs
    for k in outs[True]:
        a, b = nump(outs[True][k]), nump(outs[False][k])
        if k in ("feat", "feat_global"):
            assert np.array_equal(a, b), k
        else:
            assert_close(a, b, rel=1e-4, floor=2e-6, what=f"factored heads: {k}")


def test_face_enc_coordinate_work_ahead_is_bit_equal():
    """Face_Enc inference with the coordinate-only work (pooled clouds, level-1 / level-2 xyz kNN, nearest upsampling) on the
    forked stream equals the in-chain order bit for bit, draws the same two permutations from the CPU generator, and leaves
    the generator in the same state (gcn3d.py:241-244)."""
    from tgpose_b200.face_enc import Face_Enc
    torch.manual_seed(0)
    enc = Face_Enc().cuda().eval()
    g = torch.Generator().manual_seed(5)
    pts = torch.rand(4, 1028, 3, generator=g).cuda()
    cat = torch.randint(0, 6, (4, 1), generator=g).float().cuda()
    outs, states = [], []
    with torch.no_grad():
        for ahead in (True, False, True):
            enc.xyz_ahead = ahead
            torch.manual_seed(11)
            feat, _ = enc(pts, cat)
            torch.cuda.synchronize()
            outs.append(feat.clone())
            states.append(torch.get_rng_state().clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(states[0], states[1])


@pytest.mark.parametrize("M,K,N", [(1028, 1289, 512), (4112, 1024, 256), (300, 64, 128)])
def test_gemm_mixed_w16_operands(ops, M, K, N):
    """mixed = 2: the weight operand carries its residual in fp16 (tgp_split_mixed_w16), the third pass is fp16(a).lo16(b), and
    producers may leave the bf16(x) slot of the activation operand out (output mode 5).  Same error bound as mixed = 1 against
    an fp64 product; a mode-5 operand (its bf16 slot pre-filled with NaN patterns, never written, never read) feeds a second
    contraction."""
    g = torch.Generator().manual_seed(M + K + N + 1)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ W.double().t() + bias.double()
    out = torch.empty(M, N, device="cuda")
    ops.gemm(None, W, True, [(0, N, out, 0, 0)], bias=bias, K=K, A_split=ops.split_mixed(A), B_split=ops.split_mixed(W, w16=True),
             mixed=2)
    scale = float(ref.abs().max())
    assert float((out.double() - ref).abs().max()) <= 1.6e-5 * scale
    kp = ops.mixed_kpad(N)
    mix = torch.full((M, 2 * kp), float("nan"), device="cuda")
    ops.gemm(None, W, True, [(0, N, mix, 5, kp)], bias=bias, K=K, A_split=ops.split_mixed(A), B_split=ops.split_mixed(W, w16=True),
             mixed=2)
    m16 = mix.view(torch.int16)
    full = ops.split_mixed(out).view(torch.int16)
    assert torch.equal(m16[:, :N], full[:, :N]) and torch.equal(m16[:, 2 * kp:2 * kp + N], full[:, 2 * kp:2 * kp + N])
    fill = torch.full((1, 2 * kp), float("nan"), device="cuda").view(torch.int16)
    assert torch.equal(m16[:, kp:kp + N], fill[:, kp:kp + N].expand(M, -1))                # the bf16(x) slot was left alone
    if kp != N:
        mix.view(torch.int16)[:, N:kp] = 0
        mix.view(torch.int16)[:, 2 * kp + N:3 * kp] = 0                                     # (padding columns of the operand)
    W2 = (torch.randn(64, N, generator=g) * 0.1).cuda()
    out2 = torch.empty(M, 64, device="cuda")
    ops.gemm(None, W2, True, [(0, 64, out2, 0, 0)], K=N, A_split=mix, B_split=ops.split_mixed(W2, w16=True), mixed=2)
    ref2 = out.double() @ W2.double().t()
    assert bool(torch.isfinite(out2).all())
    assert float((out2.double() - ref2).abs().max()) <= 1.6e-5 * float(ref2.abs().max())
    # modes 4 and 5 do not mix in one launch
    with pytest.raises(RuntimeError):
        ops.gemm(None, W, True, [(0, N // 2, mix, 5, kp), (N // 2, N, mix[:, N // 4:], 4, kp)], K=K, A_split=ops.split_mixed(A),
                 B_split=ops.split_mixed(W, w16=True), mixed=2)


def test_gemm_mixed_operands_range(ops):
    """fp16(x) saturates instead of overflowing (the bf16 residual carries the remainder) and tiny values keep their relative
    accuracy through the residual: rows scaled by 1e5 / 1e-6 stay finite and accurate."""
    g = torch.Generator().manual_seed(3)
    M, K, N = 256, 320, 128
    A = torch.randn(M, K, generator=g)
    A[:64] *= 1e5          # beyond the fp16 range
    A[64:128] *= 1e-6      # fp16 subnormals / underflow
    W = torch.randn(N, K, generator=g) * 0.05
    W[:16] *= 1e-5
    A, W = A.cuda(), W.cuda()
    out = torch.empty(M, N, device="cuda")
    ops.gemm(None, W, True, [(0, N, out, 0, 0)], K=K, A_split=ops.split_mixed(A), B_split=ops.split_mixed(W), mixed=True)
    ref = A.double() @ W.double().t()
    assert bool(torch.isfinite(out).all())
    rowscale = ref.abs().amax(dim=1, keepdim=True)
    rel = ((out.double() - ref).abs() / rowscale)
    assert float(rel[128:].max()) <= 1.6e-5                     # in-range rows: full accuracy
    assert float(rel[64:128].max()) <= 2e-4                     # rows entirely below the fp16 normal range: |err| <= 2^-34 per element,
                                                                # i.e. 2^-14 relative at 1e-6 (irrelevant next to in-range terms of a sum)
    assert float(rel[:64].max()) <= 5e-3                        # saturated rows: bf16-residual accuracy, never inf
    colscale = ref[128:].abs().amax(dim=0)
    assert float(((out.double() - ref)[128:].abs() / colscale).max()) <= 2e-4   # tiny weight rows likewise


def test_posenet_inference_output_set_and_folded_ph_tail():
    """train_outputs=False (FLAGS.train = 0, PoseNet9D.py:68): only the six pose outputs are returned, they equal the
    train_outputs=True run bit for bit, and the internally computed recon (PH tail folded into one matrix) agrees with
    the unfolded chain."""
    from tgpose_b200.posenet import PoseNet9D
    g = golden("posenet")
    torch.manual_seed(0)
    net = PoseNet9D(train_outputs=True).cuda().eval()
    pts, cat = cu(g["pts"]), cu(g["cat_id"])
    with torch.no_grad():
        torch.manual_seed(7)
        full = net(pts, cat)
        recon_full = net._last_recon.clone()
        net.train_outputs = False
        torch.manual_seed(7)
        lean = net(pts, cat)
        recon_lean = net._last_recon.clone()
    assert set(lean) == {"p_green_R", "p_red_R", "f_green_R", "f_red_R", "Pred_T", "Pred_s"}
    for k in lean:
        assert torch.equal(lean[k], full[k]), k
    assert_close(nump(recon_lean), nump(recon_full), rel=1e-4, floor=2e-6, what="recon with the folded PH tail")
