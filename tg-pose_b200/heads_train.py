"""Train-mode Conv1d(k=1) + BatchNorm1d + ReLU / LeakyReLU blocks of the heads on the library's kernels.

The reference runs these as nn.Conv1d / nn.BatchNorm1d on (B, C, N) tensors (PoseR.py:26-33, PoseTs.py:31-38,
FaceRecon.py:95-117,139-141); 14.7 GFLOP per cloud forward, the dominant cost of the training step.  Here the rows
are channel-last (B*N, C):
    forward   z = x W^T + b                 tgp_gemm (tcgen05 3xTF32)
              mean, var over the rows       tgp_colsum, tgp_colsumsq_dev (two-pass)
              y = act((z - mean) * invstd * gamma + beta)   tgp_affine_act (also emits the next GEMM's split operand)
    backward  dz, dgamma, dbeta             tgp_bn_bwd
              dW = dz^T x, db = colsum(dz)  tgp_gemm_tn_tc, tgp_colsum
              dx = dz W                     tgp_gemm
BatchNorm's running statistics are updated exactly like torch's (momentum, unbiased variance, num_batches_tracked).
"""
import torch

from . import ops


class _ConvBNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, weight, bias, gamma, beta, eps, slope, x_split, want_split):
        M = x2d.shape[0]
        cout, cin = weight.shape[0], weight.shape[1]
        w2 = weight.reshape(cout, cin)
        z = torch.empty((M, cout), dtype=torch.float32, device=x2d.device)
        # forward, dx and weight-gradient contractions on MIXED operands (fp16 + bf16 cross terms, see tgp_gemm_args.mixed)
        xs = x_split if (x_split is not None and x_split.numel()) else ops.split_mixed(x2d)
        ops.gemm(None, w2, True, [(0, cout, z, 0, 0)], bias=bias, K=cin, A_split=xs, B_split=ops.split_mixed(w2.detach()),
                 mixed=True)
        mean = ops.colsum(z).view(-1) / M
        var = ops.colsumsq_dev(z, mean) / M                      # biased, as used for normalisation
        invstd = torch.rsqrt(var + eps)
        scale = gamma * invstd
        shift = beta - mean * scale
        y, y_split = ops.affine_act(z, scale, shift, slope, want_raw=True, want_split=want_split, mixed=True)
        ctx.save_for_backward(x2d, w2, z, y, mean, invstd, gamma, scale, shift, xs)     # xs: the weight gradient reads it in place
        ctx.slope = slope
        ctx.has_bias = bias is not None
        ctx.wshape = weight.shape
        if y_split is None:
            y_split = y.new_empty(0)
        ctx.mark_non_differentiable(y_split, mean, var)
        return y, y_split, mean, var

    @staticmethod
    def backward(ctx, dy, _ds, _dm, _dv):
        x2d, w2, z, y, mean, invstd, gamma, scale, shift, xs = ctx.saved_tensors
        # dz also as a mixed operand: the dX contraction AND the weight gradient (row-major operands read in place as
        # MN-major MMA operands, tgp_gemm_tn_tc_rm) consume it -- no transposing split of dz or x
        dz, dbeta, dgamma, dzm = ops.bn_bwd(dy, y, z, mean, invstd, gamma, ctx.slope, want_mixed=True,
                                             scale=scale, shift=shift)
        dw = ops.gemm_tn(dz, x2d, mixed=True, A_mixed=dzm, B_mixed=xs).view(ctx.wshape)
        db = ops.colsum(dz).view(-1) if ctx.has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            cout, cin = w2.shape
            dx = torch.empty((dz.shape[0], cin), dtype=torch.float32, device=dz.device)
            ops.gemm(None, w2, False, [(0, cin, dx, 0, 0)], K=cout, Ncols=cin, A_split=dzm if dzm is not None else ops.split_mixed(dz),
                     B_split=ops.split_mixed(w2.t().contiguous()), mixed=True)
        return dx, dw, db, dgamma, dbeta, None, None, None, None


def conv_bn_act_train(x2d, conv, bn, slope, x_split=None, want_split=False):
    """(M, Cin) channel-last rows -> act(BN_train(conv(x))) as (M, Cout) [+ its tensor-core split]; updates bn's running
    statistics like torch.nn.BatchNorm1d in train mode.  slope: 0 ReLU, 0.2 LeakyReLU(0.2), 1 no activation."""
    M = x2d.shape[0]
    y, y_split, mean, var = _ConvBNActFn.apply(x2d, conv.weight, conv.bias, bn.weight, bn.bias, bn.eps, float(slope),
                                               x_split, want_split)
    if bn.track_running_stats:
        with torch.no_grad():
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (M / max(M - 1, 1)), alpha=mom)
    return (y, y_split) if want_split else y
