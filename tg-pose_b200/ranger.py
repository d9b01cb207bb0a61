"""Ranger (RAdam + Lookahead + gradient centralisation) with the gradient clip folded in, as two kernels per step.

Mirror of tools/torch_utils/solver/ranger2020.py:44-235 (constructor arguments, state names, per-group learning
rates) for the place the reference uses it, trainer/RL_TDA.py:223-224:

    torch.nn.utils.clip_grad_norm_(self.net1.parameters(), 5)      ->  opt.clip_grad_norm_(5)
    self.optimizer.step()                                          ->  opt.step()

Layout: at construction every parameter is re-pointed into ONE flat fp32 arena (p.data becomes a view), and so are
its gradient (p.grad, a view of the gradient arena, zero-filled), exp_avg, exp_avg_sq and slow_buffer (exposed per
parameter through `state[p]` like the reference does).  One row table describes how the arenas are cut
(include/tgpose_b200.h: tgp_ranger_row); a step is tgp_ranger_reduce (row means + clip norm, one read of the
gradients) and tgp_ranger_update.  The flat gradient arena is also what the data-parallel all-reduce works on
(parallel.allreduce_flat), without the flatten / unflatten copies.

Differences from the reference, all stated:
  * `clip_grad_norm_` does not rewrite the gradients; it records max_norm and the scaling happens inside the next
    step() (the norm it returns is that of the un-clipped gradients, like torch's).
  * `zero_grad()` zero-fills the arena and keeps the views (`set_to_none` would move the gradients out of the arena).
    A gradient that autograd re-allocated elsewhere is copied back into the arena by step().
  * a parameter whose .grad is None is skipped like ranger2020.py:146-147 does, but the step counter is the
    optimiser's, not the parameter's (they coincide whenever every parameter gets a gradient on every step).
  * gc_loc=False (centralising the update instead of the gradient, ranger2020.py:217-218) runs a second, scalar
    update kernel in which a warp walks its row twice; the default gc_loc=True is the tuned path.
There is no CPU path: parameters must be CUDA tensors and the shared object must be built.
"""
import ctypes
import math

import numpy as np
import torch
from torch.optim.optimizer import Optimizer

from . import _lib

ROW_PIECE = 4096        # elements per row-table entry of a tensor that is not centralised
ALIGN = 32              # every tensor starts on a 128-byte boundary of the arenas


def radam_scalars(step, beta1, beta2, threshold):
    """ranger2020.py:183-201 -> (rectified, step_size) in Python floats, exactly the reference's expressions."""
    beta2_t = beta2 ** step
    n_sma_max = 2 / (1 - beta2) - 1
    n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
    if n_sma > threshold:
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max
                              / (n_sma_max - 2)) / (1 - beta1 ** step)
        return True, step_size
    return False, 1.0 / (1 - beta1 ** step)


def build_row_table(shapes, use_gc=True, gc_conv_only=False):
    """shapes -> (offsets, total elements, rows) with rows an (R, 4) int64 array [off, len, gc, tensor].
    centralized_gradient (ranger2020.py:31-41): tensors with more than 1 dim (more than 3 with gc_conv_only) lose the
    mean over dims 1.. per dim-0 slice -> one row per slice; the rest is cut into ROW_PIECE-element rows."""
    offsets, rows, off = [], [], 0
    for t, shape in enumerate(shapes):
        numel = int(np.prod(shape)) if len(shape) else 1
        offsets.append(off)
        gc = use_gc and len(shape) > (3 if gc_conv_only else 1)
        if numel == 0:
            pass
        elif gc:
            rl = numel // shape[0]
            rows.extend([off + r * rl, rl, 1, t] for r in range(shape[0]))
        else:
            rows.extend([off + s, min(ROW_PIECE, numel - s), 0, t] for s in range(0, numel, ROW_PIECE))
        off += (numel + ALIGN - 1) // ALIGN * ALIGN
    return offsets, off, np.asarray(rows, np.int64).reshape(-1, 4)


def _pack_rows(rows):
    """(R,4) int64 -> bytes of tgp_ranger_row[R] {long long off; int len, gc, tensor, pad}."""
    rec = np.zeros(len(rows), dtype=np.dtype([("off", "<i8"), ("len", "<i4"), ("gc", "<i4"), ("tensor", "<i4"), ("pad", "<i4")]))
    rec["off"], rec["len"], rec["gc"], rec["tensor"] = rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3]
    return rec


class Ranger(Optimizer):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(0.95, 0.999), eps=1e-5,
                 weight_decay=0, use_gc=True, gc_conv_only=False, gc_loc=True):
        # parameter checks of ranger2020.py:80-88
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f"Invalid slow update rate: {alpha}")
        if not 1 <= k:
            raise ValueError(f"Invalid lookahead steps: {k}")
        if not lr > 0:
            raise ValueError(f"Invalid Learning Rate: {lr}")
        if not eps > 0:
            raise ValueError(f"Invalid eps: {eps}")
        defaults = dict(lr=lr, alpha=alpha, k=k, step_counter=0, betas=betas, N_sma_threshhold=N_sma_threshhold,
                        eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold = N_sma_threshhold
        self.alpha, self.k = alpha, k
        self.gc_loc, self.use_gc, self.gc_conv_only = gc_loc, use_gc, gc_conv_only
        self.steps = 0
        self._max_norm = 0.0
        self._flatten()

    # ------------------------------------------------------------------ arenas
    def _flatten(self):
        self._lib = _lib.load()                       # raises when the shared object is missing: no fallback
        plist = [p for g in self.param_groups for p in g["params"]]
        if not plist:
            raise ValueError("Ranger: no parameters")
        dev = plist[0].device
        for p in plist:
            if not p.is_cuda or p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("Ranger: every parameter must be a float32 CUDA tensor on one device "
                                   "(there is no CPU path)")
        self._plist = plist
        offsets, total, rows = build_row_table([tuple(p.shape) for p in plist], self.use_gc, self.gc_conv_only)
        self._offsets, self._total = offsets, total
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._views = []
        with torch.no_grad():
            for p, off in zip(plist, offsets):
                n = p.numel()
                view = self.flat_params[off:off + n].view(p.shape)
                view.copy_(p.data)
                old_grad = p.grad
                p.data = view
                gview = self.flat_grads[off:off + n].view(p.shape)
                if old_grad is not None:
                    gview.copy_(old_grad)
                p.grad = gview
                self._views.append(gview)
        # state["slow_buffer"].copy_(p.data) happens LAZILY, at a parameter's first step() (ranger2020.py:158-168), not
        # here: the reference workflow builds the optimiser first and loads the pretrained weights afterwards
        # (engine/train.py:52-55), so the snapshot must see the loaded weights.
        self.flat_slow = torch.zeros_like(self.flat_params)
        self._slow_ready = False
        for p, off in zip(plist, offsets):
            self.state[p] = {"step": 0, **self._state_views(p, off)}
        # row ranges per parameter group (rows are emitted in parameter order)
        self._rows_host = rows
        self._rows_dev = torch.from_numpy(_pack_rows(rows).view(np.uint8).copy()).to(dev)
        self._group_rows, t0 = [], 0
        for g in self.param_groups:
            t1 = t0 + len(g["params"])
            sel = np.nonzero((rows[:, 3] >= t0) & (rows[:, 3] < t1))[0]
            self._group_rows.append((int(sel[0]), int(sel[-1]) + 1) if len(sel) else (0, 0))
            t0 = t1
        self._row_sum = torch.empty(len(rows), dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._active_key = None
        self._active_dev = None

    @property
    def n_rows(self):
        return len(self._rows_host)

    # ------------------------------------------------------------------ the reference's call sites
    def zero_grad(self, set_to_none=False):
        """zero-fill the gradient arena; the per-parameter views stay in place (see the module docstring)."""
        self.flat_grads.zero_()
        for p, v in zip(self._plist, self._views):
            if p.grad is not v:
                p.grad = v

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm):
        """torch.nn.utils.clip_grad_norm_(params, max_norm) (trainer/RL_TDA.py:223): returns the total 2-norm of the
        gradients (device scalar); the scaling by min(1, max_norm / (norm + 1e-6)) is applied inside the next step()."""
        self._max_norm = float(max_norm)
        self._reduce()
        return torch.sqrt(self._sumsq).to(torch.float32).squeeze(0)

    def _sync_grads(self):
        """gradients that are not the arena views (re-allocated by autograd after a set_to_none) are copied back;
        parameters without a gradient are marked inactive for this step."""
        active = []
        for p, v in zip(self._plist, self._views):
            g = p.grad
            if g is None:
                active.append(0)
                continue
            active.append(1)
            if g is not v and g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v
        key = tuple(active)
        if key != self._active_key:
            self._active_key = key
            self._active_dev = None if all(active) else torch.tensor(active, dtype=torch.int32, device=self.flat_params.device)
        self._reduced = False

    def _reduce(self):
        self._sync_grads()
        st = torch.cuda.current_stream(self.flat_params.device).cuda_stream
        a = self._active_dev.data_ptr() if self._active_dev is not None else None
        _lib.check(self._lib.tgp_ranger_reduce(self.flat_grads.data_ptr(), self._rows_dev.data_ptr(), self.n_rows, a,
                                               self._row_sum.data_ptr(), self._sumsq.data_ptr(), st), "tgp_ranger_reduce")
        self._reduced = True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None                                   # the reference ignores the closure too (ranger2020.py:134-140)
        if not getattr(self, "_reduced", False):
            self._reduce()
        if not self._slow_ready:                      # first step without a loaded state: slow_buffer <- p.data as it is NOW
            self.flat_slow.copy_(self.flat_params)
            self._slow_ready = True
        self.steps += 1
        st = torch.cuda.current_stream(self.flat_params.device).cuda_stream
        a = self._active_dev.data_ptr() if self._active_dev is not None else None
        for group, (r0, r1) in zip(self.param_groups, self._group_rows):
            if r1 <= r0:
                continue
            beta1, beta2 = group["betas"]
            rect, step_size = radam_scalars(self.steps, beta1, beta2, self.N_sma_threshhold)
            h = _lib.RangerHyper(beta1, beta2, group["eps"], group["weight_decay"], 1 - beta1, 1 - beta2,
                                 -step_size * group["lr"],
                                 1 if rect else 0, 1 if self.steps % group["k"] == 0 else 0, self.alpha, self._max_norm,
                                 0 if self.gc_loc else 1)
            _lib.check(self._lib.tgp_ranger_update(
                self.flat_params.data_ptr(), self.flat_grads.data_ptr(), self.flat_exp_avg.data_ptr(),
                self.flat_exp_avg_sq.data_ptr(), self.flat_slow.data_ptr(), self._rows_dev.data_ptr(), r0, r1, a,
                self._row_sum.data_ptr(), self._sumsq.data_ptr(), ctypes.byref(h), self._norm.data_ptr(), st),
                "tgp_ranger_update")
        for i, p in enumerate(self._plist):
            if self._active_key[i]:
                self.state[p]["step"] = self.steps
        try:                                          # the kernels wrote through raw pointers: tell autograd / weight caches
            torch.autograd.graph.increment_version(self._plist)
        except TypeError:
            for p in self._plist:
                torch.autograd.graph.increment_version(p)
        self._max_norm = 0.0
        self._reduced = False
        return loss

    def _state_views(self, p, off):
        n = p.numel()
        return {"exp_avg": self.flat_exp_avg[off:off + n].view(p.shape),
                "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view(p.shape),
                "slow_buffer": self.flat_slow[off:off + n].view(p.shape)}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        """checkpoint resume (trainer/RL_TDA.py:95): the loaded moments / slow buffers are copied INTO the arenas and
        state[p] is re-pointed at the arena views; the step counter resumes from the checkpoint."""
        super().load_state_dict(state_dict)
        steps = 0
        for p, off in zip(self._plist, self._offsets):
            loaded = self.state.get(p, {})
            views = self._state_views(p, off)
            for name, v in views.items():
                if name in loaded:
                    v.copy_(loaded[name])
            step = int(loaded.get("step", 0))
            steps = max(steps, step)
            self.state[p] = {"step": step, **views}
        self.steps = steps
        self._slow_ready = steps > 0                  # a resumed run keeps the checkpoint's slow buffers

    def add_param_group(self, param_group):
        """parameters are fixed once the arenas are laid out (the constructor's own calls come before that)."""
        if hasattr(self, "_plist"):
            raise RuntimeError("Ranger.add_param_group: the flat arenas are already laid out; pass every parameter "
                               "group to the constructor")
        super().add_param_group(param_group)

    @property
    def total_norm(self):
        """device scalar: the gradient norm seen by the last step()."""
        return self._norm
