"""Drop-in for losses/chamfer3D/dist_chamfer_3D.py: chamfer_3DFunction / chamfer_3DDist with the same
4-tuple result (dist1, dist2 fp32; idx1, idx2 int32) and the same backward contract, plus the
loss tails that call it (calc_cd / calc_dcd, losses/TDA_loss_sym_recon.py:411-450,495-509).

Differences from the reference wrapper, none visible to callers: outputs are allocated on the
device (the reference builds zeros on the CPU and copies them over, dist_chamfer_3D.py:33-42),
and both directions run in one launch on the current stream."""
import torch
from torch import nn
from torch.autograd import Function

from . import chamfer_3D, ops


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        device = xyz1.device
        dist1 = torch.empty(batchsize, n, device=device, dtype=torch.float32)
        dist2 = torch.empty(batchsize, m, device=device, dtype=torch.float32)
        idx1 = torch.empty(batchsize, n, device=device, dtype=torch.int32)
        idx2 = torch.empty(batchsize, m, device=device, dtype=torch.int32)
        with torch.cuda.device(device):
            ops.chamfer_forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        with torch.cuda.device(xyz1.device):
            ops.chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2)
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous().float()
        input2 = input2.contiguous().float()
        return chamfer_3DFunction.apply(input1, input2)


def calc_cd(pred, gt, return_raw=False, separate=False):
    """losses/TDA_loss_sym_recon.py:495-509."""
    dist1, dist2, idx1, idx2 = chamfer_3DDist()(pred, gt)
    cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
    cd_t = (dist1.mean(1) + dist2.mean(1))
    if separate:
        res = [torch.cat([torch.sqrt(dist1).mean(1).unsqueeze(0), torch.sqrt(dist2).mean(1).unsqueeze(0)]),
               torch.cat([dist1.mean(1).unsqueeze(0), dist2.mean(1).unsqueeze(0)])]
    else:
        res = [cd_p, cd_t]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


class _DcdTail(torch.autograd.Function):
    """exp(-alpha d) weighting by bincount(idx)^lambda, one kernel per call (the weights are detached in the
    reference, TDA_loss_sym_recon.py:433,438, so only d loss / d dist flows back)."""

    @staticmethod
    def forward(ctx, dist1, dist2, idx1, idx2, alpha, n_lambda, non_reg=False):
        need = any(ctx.needs_input_grad)
        loss, c1, c2 = ops.dcd(dist1.contiguous(), dist2.contiguous(), idx1, idx2, alpha, n_lambda, want_coef=need,
                                non_reg=non_reg)
        if need:
            ctx.save_for_backward(c1, c2)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        c1, c2 = ctx.saved_tensors
        g = gloss.contiguous().unsqueeze(1)
        return c1 * g, c2 * g, None, None, None, None, None


def calc_dcd(pred_recon, cate_gt, alpha=0.1, n_lambda=0.3, return_raw=False, non_reg=False):
    """losses/TDA_loss_sym_recon.py:411-450 -> per-cloud loss (B,); non_reg clamps the two point-count ratios to >= 1
    (:418-420)."""
    pred_recon = pred_recon.float()
    cate_gt = cate_gt.float()
    assert pred_recon.shape[0] == cate_gt.shape[0]
    dist1, dist2, idx1, idx2 = chamfer_3DDist()(pred_recon, cate_gt)
    loss = _DcdTail.apply(dist1, dist2, idx1, idx2, alpha, n_lambda, bool(non_reg))
    if return_raw:
        return [loss, dist1, dist2, idx1, idx2]
    return loss
