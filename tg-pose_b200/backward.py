"""Host side of the backward kernels (SURVEY 8a'): the autograd.Function.backward bodies.

The reference differentiates gcn3d.py:78-112 / :142-186 / :210-217 / :225-245 through torch autograd over the
materialised (B,N,k,S*C) tensors.  Here every step is one C-ABI launch on OUTPUT-sized tensors, driven by the
uint8 arg-max slots saved in the forward:

  ORL tail   out = act(scale * (f W2a^T + (g W2b^T)[cloud] + f + f_STE) + shift),  g = mean_n max_j f[idx_xyz]
      gz   = grad * act' * scale                              tgp_act_bwd
      d_gb = per-cloud column sums of gz                      tgp_colsum
      dg   = d_gb W2b, dW2b = d_gb^T g, dW2a = gz^T f         tgp_gemm / tgp_gemm_tn(_tc)
      d_f  = gz W2a + gz + scatter_argmax(dg / N)             tgp_gather_max_bwd + tgp_gemm (residual epilogue)
  layer conv (SURVEY 8a' bullets 1-2)                         tgp_layer_conv_bwd  -> d_support, d_directions
  projection fm @ [Wsup | Wc | W_STE^T] + b                   tgp_gemm (dfm), tgp_gemm_tn_tc (dW), tgp_colsum (db)
  surface conv                                                tgp_surface_conv_bwd -> d_directions
  Pool / upsample                                             tgp_gather_max_bwd / tgp_scatter_add_rows
"""
import torch

from . import ops


def _orl_tail_bwd(grad_out, out, post, feature, g, idx_xyz, arg_orl, conv2_w, B, N, C, gz_dst=None, df_dst=None):
    """backward of autograd._orl_tail; returns (gz, d_f, d_conv2_weight).  gz_dst / df_dst: optional (M,C)
    row-strided destinations (column blocks of the projection's gradient operand)."""
    M = B * N
    dev = grad_out.device
    go = grad_out.contiguous().view(M, C)
    scale, _shift, relu = post if post is not None else (None, None, False)
    if gz_dst is not None or relu or scale is not None:
        gz = ops.act_bwd(go, out.view(M, C) if relu else None, scale, relu, out=gz_dst)
    else:
        gz = go
    w2 = conv2_w.detach().reshape(C, 2 * C)
    d_gb = ops.colsum(gz, rows_per_group=N)                       # (B, C): gradient of the per-cloud bias
    dg = ops.matmul_kn(d_gb, w2[:, C:], tc=False)                           # (B, C)
    d_w2 = torch.empty((C, 2 * C), dtype=torch.float32, device=dev)
    ops.gemm_tn(d_gb, g, out=d_w2[:, C:], tc=False)
    ops.gemm_tn(gz, feature.view(M, C), out=d_w2[:, :C], mixed=True)      # gradients: fp16+bf16 mixed operands (1.5 passes)
    d_sc = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
    ops.gather_max_bwd(dg, idx_xyz, arg_orl, N, d_sc, per_cloud=True, scale=1.0 / N)
    if df_dst is None:
        df_dst = torch.empty((M, C), dtype=torch.float32, device=dev)
    ops.gemm(gz, w2[:, :C], False, [(0, C, df_dst, 0, 0)], res1=gz, res2=d_sc.view(M, C))
    return gz, df_dst, d_w2.view(C, 2 * C, 1)


def hs_surface_backward(ctx, grad_out):
    xyz, directions, ste_w, conv2_w, idx_xyz, arg, feature, g, arg_orl, out = ctx.saved_tensors
    k, S, C, post = ctx.cfg
    B, N, _ = xyz.shape
    M = B * N
    gz, d_f, d_conv2 = _orl_tail_bwd(grad_out, out, post, feature, g, idx_xyz, arg_orl, conv2_w, B, N, C)
    d_ste = ops.gemm_tn(gz, xyz.view(M, 3)).view(C, 3, 1)         # f_STE = xyz @ W_STE^T
    d_dir = ops.surface_conv_bwd(xyz, idx_xyz, directions, arg, d_f, S, C)
    return None, d_dir, d_ste, d_conv2, None, None, None, None, None, None, None


def hs_layer_backward(ctx, grad_out):
    from .autograd import _pack_layer
    (fm, weights, bias, directions, ste_w, conv2_w, rec, slab, arg, idx_xyz, feature, g, arg_orl, out) = ctx.saved_tensors
    k, S, C, post = ctx.cfg
    B, N, cin = fm.shape
    M = B * N
    SC = S * C
    # gradient operand of the packed projection, columns [d_support (slab order) | d_centre | d_f_STE] (autograd._pack_layer)
    dP = torch.empty((M, (S + 2) * C), dtype=torch.float32, device=fm.device)
    _, _, d_conv2 = _orl_tail_bwd(grad_out, out, post, feature, g, idx_xyz, arg_orl, conv2_w, B, N, C,
                                  gz_dst=dP[:, C + SC:], df_dst=dP[:, SC:SC + C])
    d_dir = ops.layer_conv_bwd(rec, directions, slab, arg, dP[:, SC:SC + C], B, N, S, C, d_support=dP[:, :SC])
    wcat, _bcat, _ws = _pack_layer(weights, bias, ste_w, S, C)    # (cin, (S+2)C)
    # dP @ wcat^T on mixed operands (gradients feed no neighbour search: fp16+bf16 accuracy is enough, 1.5 instead of 3 passes)
    d_fm = torch.empty((M, cin), dtype=torch.float32, device=fm.device)
    dPm = ops.split_mixed(dP)                                     # one mixed split of dP serves both contractions
    ops.gemm(None, wcat, True, [(0, cin, d_fm, 0, 0)], K=dP.shape[1], A_split=dPm,
             B_split=ops.split_mixed(wcat), mixed=True)
    d_wcat = ops.gemm_tn(fm.view(M, cin), dP, mixed=True, B_mixed=dPm)   # (cin, (S+2)C); row-major operands read in place
    d_bcat = ops.colsum(dP[:, :C + SC]).view(-1)
    d_weights = torch.cat([d_wcat[:, SC:SC + C],
                           d_wcat[:, :SC].reshape(cin, C // 4, S, 4).permute(0, 2, 1, 3).reshape(cin, SC)], dim=1)
    d_bias = torch.cat([d_bcat[SC:], d_bcat[:SC].reshape(C // 4, S, 4).permute(1, 0, 2).reshape(SC)])
    d_ste = d_wcat[:, C + SC:].t().reshape(C, cin, 1)
    return (None, d_fm.view(B, N, cin), d_weights, d_bias, d_dir, d_ste, d_conv2,
            None, None, None, None, None, None, None, None, None)


def pool_backward(ctx, g_pooled):
    idx_xyz, rows, arg = ctx.saved_tensors
    B, N, C = ctx.shape
    d_fm = torch.zeros((B, N, C), dtype=torch.float32, device=g_pooled.device)
    ops.gather_max_bwd(g_pooled, idx_xyz, arg, N, d_fm, rows=rows)
    return None, d_fm, None, None, None, None
