"""torch.autograd glue for the module-level forwards (kernels do the work, autograd only routes).

Forward pipelines (each step one C-ABI launch):
  HSSurfaceFn : gcn3d.py:78-112   STE gemm | xyz kNN | surface conv | ORL gather-max-mean | ORL gemm
  HSLayerFn   : gcn3d.py:142-186  fused projection+STE gemm (support slab | centre | f_STE) |
                                  feature kNN | edge records | layer conv | xyz kNN | ORL | ORL gemm
  PoolFn      : gcn3d.py:225-245  xyz kNN | gather-max on the sampled rows | row select
The backward contract is SURVEY 8(a'); see backward.py for the kernels' host side.
"""
import torch

from . import ops

def _pack_layer(weights, bias, ste_w, S, C):
    """[support in slab column order (cgroup, s, c4) | centre | STE^T] as one (in, (S+2)*C) operand, its bias and
    its tensor-core split.  The slab columns come FIRST: at S = 7 a channel group is 28 columns and S*C is a multiple of
    224, so the projection runs on 224-column tiles that each hold 8 whole channel groups and stores every (group, 32 rows)
    block with one bulk copy (csrc/gemm_tc.cu).  Cached per parameter object/version (weights are constants in inference)."""

    def build():
        cin = weights.shape[0]
        with torch.no_grad():
            w = weights.detach()
            sup = w[:, C:].reshape(cin, S, C // 4, 4).permute(0, 2, 1, 3).reshape(cin, S * C)
            wcat = torch.cat([sup, w[:, :C], ste_w.detach().reshape(C, cin).t()], dim=1).contiguous()
            b = bias.detach()
            bcat = torch.cat([b[C:].reshape(S, C // 4, 4).permute(1, 0, 2).reshape(-1), b[:C],
                              torch.zeros(C, device=b.device, dtype=b.dtype)]).contiguous()
            wsplit = ops.split_tf32(wcat, src_is_kn=True) if wcat.is_cuda else None
        return wcat, bcat, wsplit

    return ops.PARAM_CACHE.get((weights, bias, ste_w), ("pack_layer", S, C), build)


def _orl_tail(feature, idx_xyz, conv2_w, f_ste, B, N, C, post, want_arg=False, feature_split=None,
              want_split=False):
    """ORL_forward (gcn3d.py:108-112,182-186) + the STE skip: conv2(cat[f, g]) + f + f_STE, where
    conv2(cat[f, g.repeat]) = f @ W[:, :C]^T + (g @ W[:, C:]^T) broadcast over the cloud."""
    M = B * N
    w2 = conv2_w.reshape(C, 2 * C)
    if want_arg:
        g, arg = ops.orl_global(feature, idx_xyz, want_arg=True)
    else:
        g, arg = ops.orl_global(feature, idx_xyz), None
    gb = ops.linear_nk(g, w2[:, C:], tc=False)      # one row per cloud: always the exact-fp32 kernel, whatever B
    out = torch.empty((B, N, C), dtype=torch.float32, device=feature.device)
    scale, shift, relu = post if post is not None else (None, None, False)
    f2 = feature.view(M, C)
    segs = [(0, C, out.view(M, C), 0, 0)]
    out_split = None
    if want_split:   # the next layer's projection reads this result as a tensor-core operand
        out_split = ops._split_buf(M, C, feature.device)
        segs.append((0, C, out_split, 2, ops.kpad(C)))
    w2a_split = None
    if feature_split is not None:
        w2a_split = ops.PARAM_CACHE.get((conv2_w,), "orl_w2a", lambda: ops.split_tf32(w2[:, :C].detach()))
    ops.gemm(f2, w2[:, :C], True, segs, group_bias=gb, rows_per_group=N,
             res1=f2, res2=f_ste, scale=scale, shift=shift, relu=relu, A_split=feature_split, B_split=w2a_split)
    return out, g, arg, out_split


class HSSurfaceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, directions, ste_w, conv2_w, k, S, C, idx_xyz, post, want_split, grad_on):
        xyz = xyz.contiguous().float()
        B, N, _ = xyz.shape
        M = B * N
        train = grad_on and any(ctx.needs_input_grad)   # grad_on: torch.is_grad_enabled() at the call site (it is off in here)
        f_ste = ops.linear_nk(xyz.view(M, 3), ste_w.reshape(C, 3))
        if idx_xyz is None:
            idx_xyz = ops.knn_xyz(xyz, k, want64=False, want32=True)[1]
        tc = ops.tc_eligible(M, C, C)
        feature, arg, fsplit = ops.surface_conv(xyz, idx_xyz, directions, S, C, want_arg=train, want_split=tc,
                                                full=True)
        out, g, arg_orl, out_split = _orl_tail(feature, idx_xyz, conv2_w, f_ste, B, N, C, post, want_arg=train,
                                               feature_split=fsplit, want_split=want_split and tc)
        if train:
            ctx.save_for_backward(xyz, directions, ste_w, conv2_w, idx_xyz, arg, feature, g, arg_orl, out)
            ctx.cfg = (k, S, C, post)
        if out_split is None:
            out_split = out.new_empty(0)
        ctx.mark_non_differentiable(out_split)
        return out, out_split

    @staticmethod
    def backward(ctx, grad_out, _g_split):
        from . import backward as bw
        return bw.hs_surface_backward(ctx, grad_out)


class HSLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, fm, weights, bias, directions, ste_w, conv2_w, k, S, C, idx_feat, idx_xyz, post,
                fm_split, want_split, grad_on):
        xyz = xyz.contiguous().float()
        fm = fm.contiguous().float()
        B, N, cin = fm.shape
        M = B * N
        train = grad_on and any(ctx.needs_input_grad)   # grad_on: torch.is_grad_enabled() at the call site (it is off in here)
        wcat, bcat, wcat_split = _pack_layer(weights, bias, ste_w, S, C)
        dev = fm.device
        centre = torch.empty((M, C), dtype=torch.float32, device=dev)
        slab = torch.empty((C // 4, M, S * 4), dtype=torch.float32, device=dev)
        f_ste = torch.empty((M, C), dtype=torch.float32, device=dev)
        tc = ops.tc_eligible(M, C, C)
        ops.gemm(fm.view(M, cin), wcat, False,
                 [(0, S * C, slab, 1, S * 4), (S * C, S * C + C, centre, 0, 0), (S * C + C, (S + 2) * C, f_ste, 0, 0)],
                 bias=bcat, A_split=fm_split if (fm_split is not None and fm_split.numel()) else None,
                 B_split=wcat_split)
        if idx_feat is None:
            idx_feat = ops.knn_feat(fm, k, want64=False, want32=True)[1]
        rec = ops.edge_records(xyz, idx_feat)
        feature, arg, fsplit = ops.layer_conv(rec, directions, centre, slab, B, N, S, C, want_arg=train,
                                              want_split=tc, full=True)
        if idx_xyz is None:
            idx_xyz = ops.knn_xyz(xyz, k, want64=False, want32=True)[1]
        out, g, arg_orl, out_split = _orl_tail(feature, idx_xyz, conv2_w, f_ste, B, N, C, post, want_arg=train,
                                               feature_split=fsplit, want_split=want_split and tc)
        if train:
            ctx.save_for_backward(fm, weights, bias, directions, ste_w, conv2_w, rec, slab, arg, idx_xyz,
                                  feature, g, arg_orl, out)
            ctx.cfg = (k, S, C, post)
        if out_split is None:
            out_split = out.new_empty(0)
        ctx.mark_non_differentiable(out_split)
        return out, out_split

    @staticmethod
    def backward(ctx, grad_out, _g_split):
        from . import backward as bw
        return bw.hs_layer_backward(ctx, grad_out)


class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, fm, sample_idx, k, idx_xyz, grad_on):
        xyz = xyz.contiguous().float()
        fm = fm.contiguous().float()
        train = grad_on and any(ctx.needs_input_grad)   # grad_on: torch.is_grad_enabled() at the call site (it is off in here)
        if idx_xyz is None:
            idx_xyz = ops.knn_xyz(xyz, k, want64=False, want32=True)[1]
        rows = sample_idx.to(xyz.device, non_blocking=True)
        if train:
            pooled, arg = ops.gather_max(fm, idx_xyz, rows=rows, want_arg=True)
            ctx.save_for_backward(idx_xyz, rows, arg)
            ctx.shape = fm.shape
        else:
            pooled = ops.gather_max(fm, idx_xyz, rows=rows)
        v_pool = ops.select_rows(xyz, rows)
        ctx.mark_non_differentiable(v_pool)
        return v_pool, pooled

    @staticmethod
    def backward(ctx, g_v, g_pooled):
        from . import backward as bw
        return bw.pool_backward(ctx, g_pooled)


class GatherRowsFn(torch.autograd.Function):
    """indexing_neighbor_new (gcn3d.py:38-46) with its scatter-add backward (the nearest upsampling of
    FaceRecon.py:71-73 in training)."""

    @staticmethod
    def forward(ctx, tensor, index):
        ctx.save_for_backward(index)
        ctx.n = tensor.shape[1]
        ctx.tshape = tensor.shape
        return ops.gather_rows(tensor, index)

    @staticmethod
    def backward(ctx, grad):
        (index,) = ctx.saved_tensors
        return ops.scatter_add_rows(grad, index, ctx.n).view(ctx.tshape), None
