"""ctypes front-end of oracle.c + numpy compositions of the module-level forwards.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product (tg-pose_b200/)
never imports this.  Pinned against the unmodified reference by
tests/test_oracle_golden.py (fixtures: tests/golden/, made by make_golden.py).

numpy in, numpy out; all float32 C-contiguous, indices int64 (chamfer: int32),
mirroring the dtypes the reference produces.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle.c (gcc via oracle/Makefile). Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE], env={**os.environ, "CC": "gcc"})
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- primitive ops
def knn_xyz(x, k, return_dist=False):
    """gcn3d.py:14-23 for (B,N,3) inputs -> (B,N,k) int64 (bit-exact recipe, ties -> lower index)."""
    x = _f32(x)
    B, N, D = x.shape
    assert D == 3
    idx = np.empty((B, N, k), np.int64)
    dist = np.empty((B, N, N), np.float32) if return_dist else None
    lib().orc_knn_xyz(_p(x), B, N, k, _p(idx), _p(dist))
    return (idx, dist) if return_dist else idx


def knn_feat(x, k, return_dist=False):
    """gcn3d.py:14-23 for (B,N,D) feature inputs (RF-F)."""
    x = _f32(x)
    B, N, D = x.shape
    idx = np.empty((B, N, k), np.int64)
    dist = np.empty((B, N, N), np.float32) if return_dist else None
    lib().orc_knn_feat(_p(x), B, N, D, k, _p(idx), _p(dist))
    return (idx, dist) if return_dist else idx


def get_neighbor_index(x, k):
    return knn_xyz(x, k) if x.shape[-1] == 3 else knn_feat(x, k)


def get_nearest_index(target, source):
    """gcn3d.py:26-35 -> (B,N,1) int64."""
    t, s = _f32(target), _f32(source)
    B, N, _ = t.shape
    M = s.shape[1]
    idx = np.empty((B, N, 1), np.int64)
    lib().orc_nearest(_p(t), _p(s), B, N, M, _p(idx))
    return idx


def indexing_neighbor(tensor, index):
    """gcn3d.py:38-46 -> (B,M,k,C)."""
    t, index = _f32(tensor), _i64(index)
    B, N, C = t.shape
    _, M, k = index.shape
    out = np.empty((B, M, k, C), np.float32)
    lib().orc_gather(_p(t), _p(index), B, N, M, k, C, _p(out))
    return out


def direction_norm(xyz, idx):
    """gcn3d.py:48-58 -> (B,N,k,3)."""
    xyz, idx = _f32(xyz), _i64(idx)
    B, N, k = idx.shape
    out = np.empty((B, N, k, 3), np.float32)
    lib().orc_dirnorm(_p(xyz), _p(idx), B, N, k, _p(out))
    return out


def surface_conv(xyz, idx, directions, S, C):
    """gcn3d.py:91-106 -> (B,N,C)."""
    xyz, idx, directions = _f32(xyz), _i64(idx), _f32(directions)
    B, N, k = idx.shape
    out = np.empty((B, N, C), np.float32)
    lib().orc_surface_conv(_p(xyz), _p(idx), _p(directions), B, N, k, S, C, _p(out))
    return out


def layer_conv(xyz, idx, directions, P, S, C):
    """gcn3d.py:157-180 given P = fm @ weights + bias, (B,N,(S+1)C) -> (B,N,C)."""
    xyz, idx, directions, P = _f32(xyz), _i64(idx), _f32(directions), _f32(P)
    B, N, k = idx.shape
    out = np.empty((B, N, C), np.float32)
    lib().orc_layer_conv(_p(xyz), _p(idx), _p(directions), _p(P), B, N, k, S, C, _p(out))
    return out


# bench.py's CPU baseline sets this: the dense contractions then go through numpy's BLAS sgemm (what the reference's
# torch CPU path uses, MKL/OpenBLAS) instead of the plain C loop nest, so that the timed port is not handicapped.
USE_BLAS = False


def gemm_bias(A, W, bias=None):
    """A (..., K) @ W (K, Nout) + bias."""
    A, W = _f32(A), _f32(W)
    if USE_BLAS:
        out = A.reshape(-1, W.shape[0]) @ W
        if bias is not None:
            out += _f32(bias)
        return out.reshape(A.shape[:-1] + (W.shape[1],))
    K, Nout = W.shape
    M = A.size // K
    b = _f32(bias) if bias is not None else None
    out = np.empty(A.shape[:-1] + (Nout,), np.float32)
    lib().orc_gemm_bias(_p(A), _p(W), _p(b), ctypes.c_long(M), K, Nout, _p(out))
    return out


def gather_max(f, idx, rows=None, return_arg=False):
    """max over the k gathered rows; rows selects a subset of points (Pool, gcn3d.py:236-244)."""
    f, idx = _f32(f), _i64(idx)
    B, N, C = f.shape
    k = idx.shape[2]
    r = _i64(rows) if rows is not None else None
    M = N if r is None else r.shape[0]
    out = np.empty((B, M, C), np.float32)
    arg = np.empty((B, M, C), np.uint8) if return_arg else None
    lib().orc_gather_max(_p(f), _p(idx), _p(r), B, N, M, k, C, _p(out), _p(arg))
    return (out, arg) if return_arg else out


def orl_global(f, idx):
    """gcn3d.py:210-217 before the .repeat -> (B,C)."""
    f, idx = _f32(f), _i64(idx)
    B, N, C = f.shape
    g = np.empty((B, C), np.float32)
    lib().orc_orl_global(_p(f), _p(idx), B, N, idx.shape[2], C, _p(g))
    return g


def chamfer_forward(xyz1, xyz2, contract=True):
    """chamfer3D.cu:12-154 -> dist1 (B,n), dist2 (B,m) f32, idx1, idx2 int32."""
    a, b = _f32(xyz1), _f32(xyz2)
    B, n, _ = a.shape
    m = b.shape[1]
    d1, d2 = np.empty((B, n), np.float32), np.empty((B, m), np.float32)
    i1, i2 = np.empty((B, n), np.int32), np.empty((B, m), np.int32)
    c = 1 if contract else 0
    lib().orc_chamfer_nn(_p(a), _p(b), B, n, m, c, _p(d1), _p(i1))
    lib().orc_chamfer_nn(_p(b), _p(a), B, m, n, c, _p(d2), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, gd1, gd2, idx1, idx2):
    """chamfer3D.cu:155-195 -> gradxyz1 (B,n,3), gradxyz2 (B,m,3)."""
    a, b = _f32(xyz1), _f32(xyz2)
    B, n, _ = a.shape
    m = b.shape[1]
    g1, g2 = np.zeros((B, n, 3), np.float32), np.zeros((B, m, 3), np.float32)
    i1 = np.ascontiguousarray(idx1, np.int32)
    i2 = np.ascontiguousarray(idx2, np.int32)
    lib().orc_chamfer_bwd(_p(a), _p(b), _p(_f32(gd1)), _p(_f32(gd2)), _p(i1), _p(i2), B, n, m, _p(g1), _p(g2))
    return g1, g2


# ----------------------------------------------------------------------------- loss tail
def calc_cd(d1, d2):
    """TDA_loss_sym_recon.py:495-509: cd_p, cd_t from the two distance arrays."""
    cd_p = (np.sqrt(d1).mean(1) + np.sqrt(d2).mean(1)) / 2
    cd_t = d1.mean(1) + d2.mean(1)
    return cd_p.astype(np.float32), cd_t.astype(np.float32)


def calc_dcd(d1, d2, i1, i2, alpha=70.0, n_lambda=0.3, non_reg=False):
    """TDA_loss_sym_recon.py:411-450: per-cloud density-aware chamfer loss (non_reg: ratios clamped to >= 1, :418-420)."""
    B, n = d1.shape
    m = d2.shape[1]
    frac_12, frac_21 = n / m, m / n
    if non_reg:
        frac_12, frac_21 = max(1, frac_12), max(1, frac_21)
    out = np.empty(B, np.float32)
    for b in range(B):
        c1 = np.bincount(i1[b], minlength=m)
        w1 = (c1[i1[b]].astype(np.float32) ** np.float32(n_lambda) + np.float32(1e-6)) ** -1 * np.float32(frac_21)
        l1 = (-np.exp(-d1[b] * np.float32(alpha)) * w1 + 1.0).mean()
        c2 = np.bincount(i2[b], minlength=n)
        w2 = (c2[i2[b]].astype(np.float32) ** np.float32(n_lambda) + np.float32(1e-6)) ** -1 * np.float32(frac_12)
        l2 = (-np.exp(-d2[b] * np.float32(alpha)) * w2 + 1.0).mean()
        out[b] = l1 + 0.5 * l2
    return out


# ----------------------------------------------------------------------------- module forwards
def conv1x1(x, weight):
    """nn.Conv1d(kernel_size=1, bias=False) applied on the channel-last view: x (B,N,Cin), weight (Cout,Cin,1)."""
    w = _f32(weight).reshape(weight.shape[0], weight.shape[1])
    return gemm_bias(x, np.ascontiguousarray(w.T))


def orl_forward(feature, xyz, k, conv2_weight, idx_xyz=None):
    """gcn3d.py:108-112 / :182-186: conv2(cat[f, g.repeat]) + f."""
    if idx_xyz is None:
        idx_xyz = knn_xyz(xyz, k)
    g = orl_global(feature, idx_xyz)
    B, N, C = feature.shape
    cat = np.concatenate([feature, np.broadcast_to(g[:, None, :], (B, N, C))], axis=-1)
    return conv1x1(cat, conv2_weight) + feature


def hs_surface_forward(p, xyz, k, idx=None, idx_orl=None):
    """HSlayer_surface.forward, gcn3d.py:78-89.  p: dict of numpy params (state_dict names)."""
    SC = p["directions"].shape[1]
    C = p["STE_layer.weight"].shape[0]
    S = SC // C
    f_ste = conv1x1(xyz, p["STE_layer.weight"])
    if idx is None:
        idx = knn_xyz(xyz, k)
    f = surface_conv(xyz, idx, p["directions"], S, C)
    f = orl_forward(f, xyz, k, p["conv2.weight"], idx_orl if idx_orl is not None else idx)
    return f + f_ste


def hs_layer_forward(p, xyz, fm, k, idx=None, idx_orl=None):
    """HS_layer.forward, gcn3d.py:142-155."""
    C = p["STE_layer.weight"].shape[0]
    S = p["directions"].shape[1] // C
    f_ste = conv1x1(fm, p["STE_layer.weight"])
    if idx is None:
        idx = knn_feat(fm, k)
    P = gemm_bias(fm, p["weights"], p["bias"])
    f = layer_conv(xyz, idx, p["directions"], P, S, C)
    f = orl_forward(f, xyz, k, p["conv2.weight"], idx_orl)
    return f + f_ste


def pool_forward(xyz, fm, sample_idx, k=4, idx=None):
    """Pool_layer.forward, gcn3d.py:225-245; sample_idx = torch.randperm(N)[:N//rate] drawn by the caller."""
    if idx is None:
        idx = knn_xyz(xyz, k)
    pooled = gather_max(fm, idx, rows=sample_idx)
    return np.ascontiguousarray(xyz[:, sample_idx, :]), pooled


def bn_eval_relu(x, p, prefix, relu=True, eps=1e-5):
    """BatchNorm1d in eval mode on channel-last x, then ReLU (FaceRecon.py:58-65)."""
    w, b = p[prefix + ".weight"], p[prefix + ".bias"]
    mu, var = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    y = (x - mu) / np.sqrt(var + np.float32(eps)) * w + b
    y = y.astype(np.float32)
    return np.maximum(y, 0) if relu else y


def _sub(p, prefix):
    return {k[len(prefix) + 1:]: v for k, v in p.items() if k.startswith(prefix + ".")}


def face_enc_forward(p, xyz, cat_id, perm1, perm2, k=20, obj_c=6, inject=None):
    """Face_Enc.forward (eval), FaceRecon.py:39-86 -> feat (B,N0,1286).

    perm1/perm2: the two torch.randperm draws Pool_layer makes (gcn3d.py:242).
    inject: optional list of index arrays in reference call order (12 kNN + 2 nearest)
    to replay instead of computing (parity tier T2, SURVEY 8c').
    """
    it = iter(inject) if inject is not None else None

    def nxt(fn):
        return _i64(next(it)) if it is not None else fn()

    xyz = _f32(xyz)
    B, N0, _ = xyz.shape
    # conv_0: RF-P idx, ORL idx
    i0 = nxt(lambda: knn_xyz(xyz, k))
    i0o = nxt(lambda: knn_xyz(xyz, k))
    fm0 = np.maximum(hs_surface_forward(_sub(p, "conv_0"), xyz, k, i0, i0o), 0)
    # conv_1: RF-F idx, ORL idx
    i1 = nxt(lambda: knn_feat(fm0, k))
    i1o = nxt(lambda: knn_xyz(xyz, k))
    fm1 = bn_eval_relu(hs_layer_forward(_sub(p, "conv_1"), xyz, fm0, k, i1, i1o), p, "bn1")
    ip1 = nxt(lambda: knn_xyz(xyz, 4))
    n1 = N0 // 4
    v1, fp1 = pool_forward(xyz, fm1, _i64(perm1[:n1]), 4, ip1)
    k1 = min(k, n1 // 8)
    i2 = nxt(lambda: knn_feat(fp1, k1))
    i2o = nxt(lambda: knn_xyz(v1, k1))
    fm2 = bn_eval_relu(hs_layer_forward(_sub(p, "conv_2"), v1, fp1, k1, i2, i2o), p, "bn2")
    i3 = nxt(lambda: knn_feat(fm2, k1))
    i3o = nxt(lambda: knn_xyz(v1, k1))
    fm3 = bn_eval_relu(hs_layer_forward(_sub(p, "conv_3"), v1, fm2, k1, i3, i3o), p, "bn3")
    ip2 = nxt(lambda: knn_xyz(v1, 4))
    n2 = n1 // 4
    v2, fp2 = pool_forward(v1, fm3, _i64(perm2[:n2]), 4, ip2)
    k2 = min(k, n2 // 8)
    i4 = nxt(lambda: knn_feat(fp2, k2))
    i4o = nxt(lambda: knn_xyz(v2, k2))
    fm4 = hs_layer_forward(_sub(p, "conv_4"), v2, fp2, k2, i4, i4o)
    nn1 = nxt(lambda: get_nearest_index(xyz, v1))
    nn2 = nxt(lambda: get_nearest_index(xyz, v2))
    up2 = indexing_neighbor(fm2, nn1)[:, :, 0]
    up3 = indexing_neighbor(fm3, nn1)[:, :, 0]
    up4 = indexing_neighbor(fm4, nn2)[:, :, 0]
    one_hot = np.zeros((B, obj_c), np.float32)
    one_hot[np.arange(B), np.asarray(cat_id).reshape(-1).astype(np.int64)] = 1
    one_hot = np.broadcast_to(one_hot[:, None, :], (B, N0, obj_c))
    return np.concatenate([fm0, fm1, up2, up3, up4, one_hot], axis=2).astype(np.float32)


# ----------------------------------------------------------------------------- full network (heads)
def _pw(x, p, conv, bn=None, act=None):
    """1x1 Conv1d (+ eval BatchNorm + activation) on channel-last x (B,N,Cin)."""
    w = p[conv + ".weight"]
    y = gemm_bias(x, np.ascontiguousarray(w.reshape(w.shape[0], w.shape[1]).T), p.get(conv + ".bias"))
    if bn is not None:
        y = bn_eval_relu(y, p, bn, relu=False)
    if act == "relu":
        y = np.maximum(y, 0)
    elif act == "leaky":
        y = np.where(y > 0, y, np.float32(0.2) * y)
    return y.astype(np.float32)


def _lin(x, p, name):
    w = p[name + ".weight"]
    return gemm_bias(x, np.ascontiguousarray(w.T), p.get(name + ".bias"))


def _point_head(x, p, pre):
    """PoseR.py:26-39 / PoseTs.py:31-45 in eval mode: x (B,N,f) -> (B,k)."""
    h = _pw(x, p, pre + ".conv1", pre + ".bn1", "relu")
    h = _pw(h, p, pre + ".conv2", pre + ".bn2", "relu")
    h = h.max(axis=1, keepdims=True)
    h = _pw(h, p, pre + ".conv3", pre + ".bn3", "relu")
    return _pw(h, p, pre + ".conv4")[:, 0, :]


def posenet_forward(p, points, cat_id, perm1, perm2, inject=None):
    """PoseNet9D.forward (eval mode, full output set), PoseNet9D.py:33-91 + FaceRecon.py:112-202."""
    points = _f32(points)
    mean = points.mean(axis=1, keepdims=True, dtype=np.float32)
    xyz = points - mean
    enc = {k[len("face_all.encoder."):]: v for k, v in p.items() if k.startswith("face_all.encoder.")}
    feat = face_enc_forward(enc, xyz, cat_id, perm1, perm2, inject=inject)
    # PH_Predictor (FaceRecon.py:139-167)
    f5 = _pw(feat, p, "face_all.ph_pred.conv_5.0", "face_all.ph_pred.conv_5.1", "leaky")
    pooled = f5.max(axis=1)
    fa = _lin(np.concatenate([pooled, pooled], 1), p, "face_all.ph_pred.linear1")
    fa = bn_eval_relu(fa, p, "face_all.ph_pred.bn5", relu=False)
    fa = np.where(fa > 0, fa, np.float32(0.2) * fa).astype(np.float32)
    pi1 = _lin(fa, p, "face_all.ph_pred.linear2")
    pi2 = _lin(fa, p, "face_all.ph_pred.linear3")
    h1 = 1.0 / (1.0 + np.exp(-pi1))
    h2 = 1.0 / (1.0 + np.exp(-pi2))
    feat_ph = feat + (_lin(pi1, p, "face_all.ph_pred.linear4") + _lin(pi2, p, "face_all.ph_pred.linear5"))[:, None, :]
    # Face_Dec (FaceRecon.py:112-117)
    d = "face_all.decoder."
    h = _pw(feat_ph, p, d + "conv1d_block.0", d + "conv1d_block.1", "relu")
    h = _pw(h, p, d + "conv1d_block.3", d + "conv1d_block.4", "relu")
    h = _pw(h, p, d + "conv1d_block.6", d + "conv1d_block.7", "relu")
    h = _pw(h, p, d + "recon_head.0", d + "recon_head.1", "relu")
    recon = _pw(h, p, d + "recon_head.3")
    green = _point_head(feat, p, "rot_green")
    red = _point_head(feat, p, "rot_red")
    ts = _point_head(np.concatenate([feat, xyz], axis=2), p, "ts")

    def unit(v):
        return v / (np.linalg.norm(v, axis=1, keepdims=True) + np.float32(1e-6))

    return {"recon": recon + mean, "p_green_R": unit(green[:, 1:]), "p_red_R": unit(red[:, 1:]),
            "f_green_R": 1.0 / (1.0 + np.exp(-green[:, 0])), "f_red_R": 1.0 / (1.0 + np.exp(-red[:, 0])),
            "Pred_T": ts[:, 0:3] + mean[:, 0, :], "Pred_s": ts[:, 3:6], "h1": h1, "h2": h2, "feat": feat,
            "feat_global": feat.max(axis=1)}


# ----------------------------------------------------------------------------- backward (SURVEY 8a')
# What torch autograd computes through the reference graph (gcn3d.py:78-112, :142-186, :210-217, :225-245),
# restated in closed form with float64 accumulation.  Pinned against the reference's own autograd by
# tests/golden/backward.npz (make_golden.py: backward_cases).  Small cases only (pure numpy).
def _unit_dirs(directions, S, C):
    """F.normalize(directions, dim=0) -> (u (3,S*C) float64, norm (S*C,))."""
    v = np.asarray(directions, np.float64)
    nrm = np.maximum(np.sqrt((v * v).sum(0)), 1e-12)
    return v / nrm, nrm


def _normalize_bwd(directions, du):
    """backward of F.normalize(dim=0): dv = (du - u (u.du)) / ||v||."""
    u, nrm = _unit_dirs(directions, 0, 0)
    return ((du - u * (u * du).sum(0)) / nrm).astype(np.float32)


def _theta(xyz, idx, directions, S, C):
    d = direction_norm(xyz, idx).astype(np.float64)                  # (B,N,k,3)
    u, _ = _unit_dirs(directions, S, C)
    th = np.maximum(d @ u, 0.0)                                      # (B,N,k,S*C)
    return d, th


def surface_conv_backward(xyz, idx, directions, S, C, G):
    """d directions of gcn3d.py:91-106 for upstream G (B,N,C)."""
    d, th = _theta(xyz, idx, directions, S, C)
    B, N, k, _ = th.shape
    th5 = th.reshape(B, N, k, S, C)
    jstar = th5.argmax(2)                                            # (B,N,S,C)
    a = np.asarray(G, np.float64)[:, :, None, :] / S                 # (B,N,1,C)
    thmax = np.take_along_axis(th5, jstar[:, :, None], 2)[:, :, 0]   # (B,N,S,C)
    dth = np.where(thmax > 0, a, 0.0)                                # (B,N,S,C)
    dsel = np.take_along_axis(d[:, :, :, None, None, :], jstar[:, :, None, :, :, None], 2)[:, :, 0]  # (B,N,S,C,3)
    du = (dth[..., None] * dsel).sum((0, 1)).reshape(S * C, 3).T     # (3, S*C)
    return _normalize_bwd(directions, du)


def layer_conv_backward(xyz, idx, directions, P, S, C, G):
    """(dP (B,N,(S+1)C), d directions) of gcn3d.py:157-180 for upstream G (B,N,C)."""
    idx = _i64(idx)
    d, th = _theta(xyz, idx, directions, S, C)
    B, N, k, _ = th.shape
    P64 = np.asarray(P, np.float64)
    sup = P64[..., C:]                                               # (B,N,S*C)
    gathered = np.stack([sup[b][idx[b]] for b in range(B)])          # (B,N,k,S*C)
    v5 = (th * gathered).reshape(B, N, k, S, C)
    jstar = v5.argmax(2)
    a = np.asarray(G, np.float64)[:, :, None, :] / S
    th_s = np.take_along_axis(th.reshape(B, N, k, S, C), jstar[:, :, None], 2)[:, :, 0]
    sup_s = np.take_along_axis(gathered.reshape(B, N, k, S, C), jstar[:, :, None], 2)[:, :, 0]
    dP = np.zeros_like(P64)
    dP[..., :C] = G
    nb = np.take_along_axis(idx[:, :, :, None, None], jstar[:, :, None], 2)[:, :, 0]   # (B,N,S,C) neighbour row
    contrib = a * th_s                                               # (B,N,S,C)
    cols = C + (np.arange(S)[:, None] * C + np.arange(C)[None, :])   # (S,C)
    for b in range(B):
        np.add.at(dP[b], (nb[b].reshape(-1), np.broadcast_to(cols, (N, S, C)).reshape(-1)), contrib[b].reshape(-1))
    dth = np.where(th_s > 0, a * sup_s, 0.0)
    dsel = np.take_along_axis(d[:, :, :, None, None, :], jstar[:, :, None, :, :, None], 2)[:, :, 0]
    du = (dth[..., None] * dsel).sum((0, 1)).reshape(S * C, 3).T
    return dP.astype(np.float32), _normalize_bwd(directions, du)


def gather_max_backward(f, idx, G, rows=None):
    """df of max_j f[b, idx[b, rows[m], j], c] for upstream G (B,M,C)."""
    f, idx = np.asarray(f, np.float64), _i64(idx)
    B, N, C = f.shape
    r = np.arange(N) if rows is None else _i64(rows)
    df = np.zeros_like(f)
    for b in range(B):
        gath = f[b][idx[b][r]]                                       # (M,k,C)
        j = gath.argmax(1)                                           # (M,C)
        nb = np.take_along_axis(idx[b][r][:, :, None], j[:, None, :], 1)[:, 0]   # (M,C)
        np.add.at(df[b], (nb.reshape(-1), np.broadcast_to(np.arange(C), nb.shape).reshape(-1)),
                  np.asarray(G[b], np.float64).reshape(-1))
    return df.astype(np.float32)


def _orl_backward(f, idx_xyz, conv2_weight, gz):
    """backward of conv2(cat[f, g.repeat]) + f with g = mean_n max_j f[idx_xyz] -> (d_f, d_conv2_weight)."""
    B, N, C = f.shape
    w2 = np.asarray(conv2_weight, np.float64).reshape(C, 2 * C)
    g = orl_global(f, idx_xyz).astype(np.float64)
    gz64 = np.asarray(gz, np.float64)
    d_gb = gz64.sum(1)                                               # (B,C)
    dW2a = np.einsum("bno,bni->oi", gz64, np.asarray(f, np.float64))
    dW2b = d_gb.T @ g
    dg = d_gb @ w2[:, C:]
    d_f = gz64 @ w2[:, :C] + gz64
    d_f += gather_max_backward(f, idx_xyz, np.broadcast_to((dg / N)[:, None, :], (B, N, C)))
    return d_f.astype(np.float32), np.concatenate([dW2a, dW2b], 1).reshape(C, 2 * C, 1).astype(np.float32)


def hs_surface_backward(p, xyz, k, idx, idx_orl, G):
    """parameter gradients of HSlayer_surface.forward (gcn3d.py:78-89) for upstream G."""
    C = p["STE_layer.weight"].shape[0]
    S = p["directions"].shape[1] // C
    f = surface_conv(xyz, idx, p["directions"], S, C)
    d_f, d_conv2 = _orl_backward(f, idx_orl, p["conv2.weight"], G)
    d_ste = np.einsum("bno,bni->oi", np.asarray(G, np.float64), np.asarray(xyz, np.float64))
    return {"directions": surface_conv_backward(xyz, idx, p["directions"], S, C, d_f),
            "STE_layer.weight": d_ste.reshape(C, 3, 1).astype(np.float32), "conv2.weight": d_conv2}


def hs_layer_backward(p, xyz, fm, k, idx, idx_orl, G):
    """(d_fm, parameter gradients) of HS_layer.forward (gcn3d.py:142-155) for upstream G."""
    C = p["STE_layer.weight"].shape[0]
    S = p["directions"].shape[1] // C
    cin = fm.shape[-1]
    P = gemm_bias(fm, p["weights"], p["bias"])
    f = layer_conv(xyz, idx, p["directions"], P, S, C)
    d_f, d_conv2 = _orl_backward(f, idx_orl, p["conv2.weight"], G)
    dP, d_dir = layer_conv_backward(xyz, idx, p["directions"], P, S, C, d_f)
    fm64, dP64, G64 = np.asarray(fm, np.float64), dP.astype(np.float64), np.asarray(G, np.float64)
    wste = np.asarray(p["STE_layer.weight"], np.float64).reshape(C, cin)
    d_fm = dP64 @ np.asarray(p["weights"], np.float64).T + G64 @ wste
    grads = {"weights": np.einsum("bni,bno->io", fm64, dP64).astype(np.float32),
             "bias": dP64.sum((0, 1)).astype(np.float32), "directions": d_dir,
             "STE_layer.weight": np.einsum("bno,bni->oi", G64, fm64).reshape(C, cin, 1).astype(np.float32),
             "conv2.weight": d_conv2}
    return d_fm.astype(np.float32), grads


# ----------------------------------------------------------------------------- optimiser step
def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (trainer/RL_TDA.py:223): total 2-norm over all tensors, coefficient
    max_norm / (total + 1e-6) clamped to 1.  Returns (total_norm, coefficient)."""
    total = np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads))
    return total, min(1.0, max_norm / (total + 1e-6))


def radam_scalars(step, beta1, beta2, threshold):
    """tools/torch_utils/solver/ranger2020.py:183-201 -> (N_sma > threshold, step_size), double arithmetic like the
    reference's Python floats."""
    beta2_t = beta2 ** step
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma > threshold:
        step_size = np.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) \
            / (1 - beta1 ** step)
        return True, float(step_size)
    return False, 1.0 / (1 - beta1 ** step)


def ranger_step(state, grads, lr=1e-3, alpha=0.5, k=6, threshold=5, betas=(0.95, 0.999), eps=1e-5, weight_decay=0.0,
                use_gc=True, gc_conv_only=False, max_norm=None, gc_loc=True):
    """One clip_grad_norm_ + Ranger.step() (ranger2020.py:133-235) in float32 numpy.  gc_loc=True centralises the
    gradient (:170-171), gc_loc=False the update G_grad (:217-218; in place, so on un-rectified steps -- where G_grad
    IS exp_avg -- the moment is centralised too).
    state: {"step": int, "p": [...], "m": [...], "v": [...], "slow": [...]} updated in place; returns the total norm."""
    f = np.float32
    total, coef = clip_coef(grads, max_norm) if max_norm else (0.0, 1.0)
    state["step"] += 1
    t = state["step"]
    rect, step_size = radam_scalars(t, betas[0], betas[1], threshold)
    for i, g in enumerate(grads):
        g = (g * f(coef)).astype(f) if max_norm else g.astype(f)
        gc = use_gc and g.ndim > (3 if gc_conv_only else 1)                       # centralized_gradient :31-41
        if gc and gc_loc:
            g = g - g.mean(axis=tuple(range(1, g.ndim)), keepdims=True, dtype=f)
        p, m, v = state["p"][i], state["m"][i], state["v"][i]
        v *= f(betas[1])
        v += f(1 - betas[1]) * g * g                                               # :174
        m *= f(betas[0])
        m += f(1 - betas[0]) * g                                                   # :177
        G = m / (np.sqrt(v) + f(eps)) if rect else m                               # :208-212 (`G_grad = exp_avg`: an ALIAS)
        if weight_decay != 0:
            G += f(weight_decay) * p        # :214-215, in place: on un-rectified steps this also lands in exp_avg
        if gc and not gc_loc:
            G -= G.mean(axis=tuple(range(1, G.ndim)), keepdims=True, dtype=f)      # :217-218, in place as well
        p += f(-step_size * lr) * G                                                # :220
        if t % k == 0:                                                             # :225-231
            slow = state["slow"][i]
            slow += f(alpha) * (p - slow)
            p[...] = slow
    return total


def ranger_init(params):
    """state as ranger2020.py:155-165 creates it on the first step."""
    return {"step": 0, "p": [np.array(p, np.float32) for p in params], "m": [np.zeros_like(p, np.float32) for p in params],
            "v": [np.zeros_like(p, np.float32) for p in params], "slow": [np.array(p, np.float32) for p in params]}
