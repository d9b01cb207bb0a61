"""Synthetic restatement of the RL_TDA training step (reference trainer/RL_TDA.py:110-226) on the new path.

The reference trainer cannot be imported (tools/*.py sources are missing), so this keeps the parts of
the step that exercise the hot path and restates the glue around them:

  net1(PC, obj_id)                       full PoseNet9D in train mode (RL_TDA.py:116): Face_Enc forward on the
                                         sm_100a kernels, heads as the reference's torch layers
  net2(Aug_PC, obj_id) under no_grad     PoseNet9D(only_encoder=True) on the augmented cloud (RL_TDA.py:20,117-118; the
                                         reference never switches it to eval, so it runs in train mode)
  RL loss                                feat_consistency_loss(feat_global_1, feat_global_2) (consistency_loss.py:11-16)
  recon_1 / recon consistency            prop_sym_matching_loss(PC, recon_1, R, t, sym) and (recon_1, recon_2, ...)
                                         (consistency_loss.py:19-79; RL_TDA.py:133-137, the latter weighted 0.2)
  lr schedule                            flat_and_anneal (lr_scheduler.py:177-279 with config.py:123-130), stepped after
                                         the optimiser like RL_TDA.py:224-225
  R_DCD (TDA_loss_sym_recon.py:326-343)  calc_dcd(points_re_n, prototype (B,1024,3), alpha=70, n_lambda=0.3):
                                         chamfer3D forward + the in-kernel bincount tail; here points_re_n is the
                                         network's `recon` output moved by the predicted translation and size
                                         (keeps chamfer -> recon -> decoder -> Face_Enc in the autograd graph)
  recon loss (RL_TDA.py:133-135,214)     cd_t between recon and the input cloud (calc_cd, :495-509)
  pose terms                             smooth-L1 of the rotation axes / T / s against synthetic ground truth
  total.backward()                       chamfer backward + the backward kernels of SURVEY 8a' + torch heads
  all-reduce                             NCCL gradient averaging (parallel.allreduce_gradients); world 1: no-op
  clip_grad_norm_(5); optimizer.step()   RL_TDA.py:222-224 with the reference's optimiser, Ranger
                                         (config.py:126 optimizer_type = 'Ranger'): ranger.Ranger, two kernels per
                                         step over flat arenas; optimizer="adam" keeps torch's fused Adam
"""
import math

import torch
import torch.nn.functional as F

from . import parallel
from .dist_chamfer_3D import calc_cd, calc_dcd
from .ranger import Ranger


# ------------------------------------------------------------------------------------------ loss glue (outside the hot path)
def feat_consistency_loss(x1, x2, feat_consist_w=2.0):
    """losses/consistency_loss.py:11-16 (FLAGS.feat_consist_w = 2.0, config.py:83)."""
    x1 = F.normalize(x1, dim=1)
    x2 = F.normalize(x2, dim=1)
    return feat_consist_w * (2 - 2 * (x1 * x2).sum() / x1.shape[0])


def prop_sym_matching_loss(PC, PC_re, gt_R, gt_t, sym):
    """losses/consistency_loss.py:19-79: L1 between the reconstruction and the symmetry-aware target cloud --
    the input itself (no reflection symmetry), its y-axis reflection (sym[0] = 1 and another flag set: can, bowl, bottle) or
    its yx reflection (sym[0] = 0, sym[1] = 1: laptop, mug), built in the canonical frame and moved back by (R, t); clouds
    with sym = (1,0,0,0) contribute |0 - 0|."""
    bs = PC.shape[0]
    cano = torch.bmm(gt_R.transpose(1, 2), (PC - gt_t.view(bs, 1, 3)).transpose(1, 2)).transpose(1, 2)
    rest = sym[:, 1:].sum(dim=-1)
    y_flag = ((sym[:, 0] == 1) & (rest > 0)).view(-1, 1, 1)
    yx_flag = ((sym[:, 0] == 0) & (sym[:, 1] == 1)).view(-1, 1, 1)
    no_flag = ((sym[:, 0] == 0) & (sym[:, 1] != 1)).view(-1, 1, 1)
    zero_flag = ((sym[:, 0] == 1) & (rest == 0)).view(-1, 1, 1)

    def back(sign):
        p = cano * torch.tensor(sign, dtype=cano.dtype, device=cano.device).view(1, 1, 3)
        return (torch.matmul(gt_R, p.transpose(1, 2)) + gt_t.unsqueeze(-1)).transpose(1, 2)

    zero = torch.zeros_like(PC)
    target = torch.where(yx_flag, back([1., 1., -1.]), zero) + torch.where(y_flag, back([-1., 1., -1.]), zero) \
        + torch.where(no_flag, PC, zero)
    pc_re = torch.where(zero_flag, torch.zeros_like(PC_re), PC_re)
    return F.l1_loss(target, pc_re)


def flat_and_anneal_factor(it, total_iters, warmup_iters=1000, warmup_factor=0.001, anneal_point=0.72, target_lr_factor=0.0):
    """lr factor of flat_and_anneal_lr_scheduler (lr_scheduler.py:177-279) for the flags' defaults: linear warm-up,
    flat, cosine anneal from anneal_point * total_iters (config.py:123-130)."""
    anneal_start = anneal_point * total_iters
    if it < warmup_iters:
        alpha = float(it) / warmup_iters
        return warmup_factor * (1 - alpha) + alpha
    if it >= anneal_start:
        return target_lr_factor + 0.5 * (1 - target_lr_factor) * (1 + math.cos(math.pi * ((float(it) - anneal_start) / (total_iters - anneal_start))))
    return 1.0


class FlatAndAnneal:
    """scheduler.step() of RL_TDA.py:225 for any optimiser with param_groups (torch's LambdaLR semantics: the factor of
    iteration `it` multiplies each group's initial lr)."""

    def __init__(self, optimizer, total_iters, **kw):
        self.opt, self.total, self.kw, self.it = optimizer, total_iters, kw, 0
        self.base = [g["lr"] for g in optimizer.param_groups]
        self._apply()

    def _apply(self):
        f = flat_and_anneal_factor(self.it, self.total, **self.kw)
        for g, b in zip(self.opt.param_groups, self.base):
            g["lr"] = b * f

    def step(self):
        self.it += 1
        self._apply()


def synthetic_targets(batch, seed, device):
    """prototype cloud (obj_model/points_*.npy is (1024,3)), pose ground truth (R, t, s), symmetry flags
    (datasets: sym_info (4,)) cycling through the four classes of consistency_loss.py:41-79."""
    g = torch.Generator().manual_seed(seed)
    proto = (torch.rand(batch, 1024, 3, generator=g) - 0.5) * 0.3
    q = torch.linalg.qr(torch.randn(batch, 3, 3, generator=g))[0]
    R = q * torch.sign(torch.linalg.det(q)).view(batch, 1, 1)
    t = torch.stack([torch.rand(batch, generator=g) * 0.6 - 0.3, torch.rand(batch, generator=g) * 0.6 - 0.3,
                     torch.rand(batch, generator=g) * 0.8 + 0.6], dim=1)
    s = torch.rand(batch, 3, generator=g) * 0.2 + 0.1
    classes = torch.tensor([[0, 0, 0, 0], [1, 1, 0, 0], [0, 1, 0, 0], [1, 0, 0, 0]])
    sym = classes[torch.arange(batch) % 4]
    d = {"proto": proto, "R": R, "green": R[:, :, 1].contiguous(), "red": R[:, :, 0].contiguous(), "T": t, "s": s, "sym": sym}
    return {k: v.to(device) for k, v in d.items()}


def augment(pts, seed):
    """stand-in for the dataset's aug_pcl_in (datasets/data_augmentation.py): the same cloud, jittered and slightly scaled."""
    g = torch.Generator().manual_seed(seed)
    return pts * (1.0 + 0.05 * (torch.rand(pts.shape[0], 1, 1, generator=g) - 0.5)) + 0.002 * torch.randn(pts.shape, generator=g)


def losses(out, pts, tgt, out2=None):
    """loss_dict of RL_TDA.py:121-178.  TDA_loss (losses/TDA_loss_sym_recon.py, not importable) is restated by its
    hot-path terms: R_DCD on the reconstruction moved into the prototype frame, the chamfer recon term, pose terms."""
    recon = out["recon"]
    pts_n = (recon - out["Pred_T"].unsqueeze(1)) * (1.0 + out["Pred_s"].unsqueeze(1))
    r_dcd = calc_dcd(pts_n, tgt["proto"], alpha=70, n_lambda=0.3).mean()
    _, cd_t = calc_cd(recon, pts)
    pose = (F.smooth_l1_loss(out["p_green_R"], tgt["green"]) + F.smooth_l1_loss(out["p_red_R"], tgt["red"])
            + F.smooth_l1_loss(out["Pred_T"], tgt["T"]) + F.smooth_l1_loss(out["Pred_s"], tgt["s"]))
    ls = {"R_DCD": r_dcd, "recon": cd_t.mean(), "pose": pose}
    if out2 is not None:
        ls["RL_loss"] = feat_consistency_loss(out["feat_global"], out2["feat_global"])
        ls["recon_1_loss"] = prop_sym_matching_loss(pts, recon, tgt["R"], tgt["T"], tgt["sym"])
        ls["recon_consistency_loss"] = 0.2 * prop_sym_matching_loss(recon, out2["recon"], tgt["R"], tgt["T"], tgt["sym"])
    return ls


class TrainStep:
    """one optimisation step: net1 forward, net2 forward (no grad) on the augmented cloud, losses, backward with the
    gradient all-reduce overlapped, clip, optimizer step, scheduler step."""

    def __init__(self, net, lr=1e-4, seed=7, optimizer="ranger", net2=None, total_iters=20000, overlap=True):
        self.net = net
        self.net2 = net2
        self.params = [p for p in net.parameters() if p.requires_grad]
        self.ranger = optimizer == "ranger"
        if self.ranger:
            self.opt = Ranger(self.params, lr=lr)
        elif optimizer == "adam":
            self.opt = torch.optim.Adam(self.params, lr=lr, fused=self.params[0].is_cuda)
        else:
            raise ValueError(f"TrainStep: unknown optimizer {optimizer!r}")
        # the reference seeds ONCE at start-up (engine/train.py seed_init_fn): every rank starts from the same CPU-RNG
        # state and consumes it identically (Pool_layer's randperm draws), so the ranks stay in lock-step while the
        # subsample -- and every dropout mask -- changes from step to step.
        self.seed = seed
        parallel.seed_for_forward(seed)
        self.buckets = 0
        self.ar_events = None          # bench.py sets this to a list to time the collective with CUDA events
        self.sched = FlatAndAnneal(self.opt, total_iters)
        # DDP-style overlap: the flat gradient arena is cut into ~32 MB slices; a slice's all-reduce is
        # launched from an autograd hook as soon as the last of its gradients has been accumulated, while backward goes on
        self.overlap = None
        if self.ranger and overlap and parallel.dist.is_initialized() and parallel.dist.get_world_size() > 1:
            self.overlap = parallel.OverlappedAllReduce(self.opt.flat_grads, self.params, self.opt._offsets)

    def __call__(self, pts, cat, tgt, aug_pts=None):
        self.net.train()
        out = self.net(pts, cat)
        out2 = None
        if self.net2 is not None and aug_pts is not None:
            with torch.no_grad():
                out2 = self.net2(aug_pts, cat)
        ls = losses(out, pts, tgt, out2)
        tda = ls["recon"] + ls["R_DCD"] + ls["pose"]
        if out2 is not None:                       # RL_TDA.py:214
            total = 0.1 * ls["RL_loss"] + 0.1 * ls["recon_1_loss"] + 0.1 * ls["recon_consistency_loss"] + 0.9 * tda
        else:                                      # only_TDA (RL_TDA.py:119-120,139-140)
            total = 0.9 * tda
        self.last_losses = {k: v.detach() for k, v in ls.items()}
        if self.ranger:
            self.opt.zero_grad()
            ev = None
            if self.ar_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            if self.overlap is not None:
                self.overlap.begin()
                total.backward()
                if ev is not None:
                    ev[0].record()                 # what is NOT hidden behind backward: the wait for the slices in flight
                self.buckets = self.overlap.finish()
            else:
                total.backward()
                if ev is not None:
                    ev[0].record()
                self.buckets = parallel.allreduce_flat(self.opt.flat_grads)
            if ev is not None:
                ev[1].record()
                self.ar_events.append(ev)
            self.opt.clip_grad_norm_(5.0)
            self.opt.step()
            self.sched.step()
            return total.detach()
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.buckets = parallel.allreduce_gradients(self.params)
        torch.nn.utils.clip_grad_norm_(self.params, 5.0)
        self.opt.step()
        self.sched.step()
        return total.detach()
