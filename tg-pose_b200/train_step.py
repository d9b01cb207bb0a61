"""Synthetic restatement of the RL_TDA training step (reference trainer/RL_TDA.py:110-226) on the new path.

The reference trainer cannot be imported (tools/*.py sources are missing), so this keeps the parts of
the step that exercise the hot path and states the rest as simple stand-ins:

  net1(PC, obj_id)                       full PoseNet9D in train mode (RL_TDA.py:116): Face_Enc forward on the
                                         sm_100a kernels, heads as the reference's torch layers
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
import torch
import torch.nn.functional as F

from . import parallel
from .dist_chamfer_3D import calc_cd, calc_dcd
from .ranger import Ranger


def synthetic_targets(batch, seed, device):
    """prototype cloud (obj_model/points_*.npy is (1024,3)) and pose ground truth."""
    g = torch.Generator().manual_seed(seed)
    proto = (torch.rand(batch, 1024, 3, generator=g) - 0.5) * 0.3
    axis = F.normalize(torch.randn(batch, 2, 3, generator=g), dim=-1)
    t = torch.stack([torch.rand(batch, generator=g) * 0.6 - 0.3, torch.rand(batch, generator=g) * 0.6 - 0.3,
                     torch.rand(batch, generator=g) * 0.8 + 0.6], dim=1)
    s = torch.rand(batch, 3, generator=g) * 0.2 + 0.1
    return {k: v.to(device) for k, v in {"proto": proto, "green": axis[:, 0], "red": axis[:, 1], "T": t, "s": s}.items()}


def losses(out, pts, tgt):
    recon = out["recon"]
    pts_n = (recon - out["Pred_T"].unsqueeze(1)) * (1.0 + out["Pred_s"].unsqueeze(1))
    r_dcd = calc_dcd(pts_n, tgt["proto"], alpha=70, n_lambda=0.3).mean()
    _, cd_t = calc_cd(recon, pts)
    pose = (F.smooth_l1_loss(out["p_green_R"], tgt["green"]) + F.smooth_l1_loss(out["p_red_R"], tgt["red"])
            + F.smooth_l1_loss(out["Pred_T"], tgt["T"]) + F.smooth_l1_loss(out["Pred_s"], tgt["s"]))
    return {"R_DCD": r_dcd, "recon": cd_t.mean(), "pose": pose}


class TrainStep:
    """one optimisation step: forward, loss, backward, gradient all-reduce, clip, optimizer step."""

    def __init__(self, net, lr=1e-4, seed=7, optimizer="ranger"):
        self.net = net
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

    def __call__(self, pts, cat, tgt):
        self.net.train()
        out = self.net(pts, cat)
        ls = losses(out, pts, tgt)
        total = 0.1 * ls["recon"] + 0.9 * ls["R_DCD"] + 0.1 * ls["pose"]        # weights as RL_TDA.py:214
        if self.ranger:
            self.opt.zero_grad()
            total.backward()
            ev = None
            if self.ar_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self.buckets = parallel.allreduce_flat(self.opt.flat_grads)
            if ev is not None:
                ev[1].record()
                self.ar_events.append(ev)
            self.opt.clip_grad_norm_(5.0)
            self.opt.step()
            return total.detach()
        self.opt.zero_grad(set_to_none=True)
        total.backward()
        self.buckets = parallel.allreduce_gradients(self.params)
        torch.nn.utils.clip_grad_norm_(self.params, 5.0)
        self.opt.step()
        return total.detach()
