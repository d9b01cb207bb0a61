"""GPU parity of the fused clip + Ranger step (tgp_ranger_reduce / tgp_ranger_update behind ranger.Ranger) against the
golden vectors recorded from the reference's own optimiser (tests/golden/ranger.npz: tools/torch_utils/solver/
ranger2020.py behind torch.nn.utils.clip_grad_norm_) and against the numpy oracle on shapes the golden file lacks."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from util import golden

pytestmark = pytest.mark.gpu

CONFIGS = {"default": dict(lr=1e-3),
           "wd_convonly": dict(lr=3e-3, weight_decay=0.01, gc_conv_only=True, betas=(0.9, 0.99), k=4, alpha=0.3),
           "nogc": dict(lr=1e-2, use_gc=False, eps=1e-8),
           "gc_update": dict(lr=2e-3, gc_loc=False, weight_decay=0.005)}


def nump(t):
    return t.detach().cpu().numpy()


def close_update(p, ref, p0, what):
    """parameters agree to a few fp32 ulps of the value plus 1e-4 of the distance travelled from p0."""
    tol = 4e-7 * np.maximum(np.abs(ref), 1.0) + 1e-4 * np.abs(ref - p0)
    err = np.abs(p.astype(np.float64) - ref)
    assert (err <= tol).all(), f"{what}: worst {float((err - tol).max()):.3e} over tolerance; max err {float(err.max()):.3e}"


def close_rel(a, ref, what, rel=1e-5):
    scale = float(np.abs(ref).max()) + 1e-30
    err = float(np.abs(a.astype(np.float64) - ref).max())
    assert err <= rel * scale, f"{what}: {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("tag", list(CONFIGS))
def test_ranger_vs_reference_golden(tag):
    from tgpose_b200.ranger import Ranger
    z = golden("ranger")
    n, steps, snaps = int(z["n_tensors"]), int(z["steps"]), set(int(s) for s in z["snaps"])
    params = [torch.nn.Parameter(torch.from_numpy(z[f"p0_{i}"]).cuda()) for i in range(n)]
    opt = Ranger(params, **CONFIGS[tag])
    for t in range(steps):
        opt.zero_grad()
        for i, p in enumerate(params):
            p.grad.copy_(torch.from_numpy(z[f"g_{t}_{i}"]))
        v0 = params[0]._version
        norm = float(opt.clip_grad_norm_(5))
        opt.step()
        assert params[0]._version > v0                     # weight caches keyed on _version see the update
        assert abs(norm - z[f"{tag}_norms"][t]) <= 1e-5 * norm, (t, norm, z[f"{tag}_norms"][t])
        assert abs(float(opt.total_norm) - norm) <= 1e-6 * norm
        if t in snaps:
            for i, p in enumerate(params):
                close_update(nump(p), z[f"{tag}_p_{t}_{i}"], z[f"p0_{i}"], f"{tag} step {t + 1} tensor {i}")
    for i, p in enumerate(params):
        st = opt.state[p]
        assert st["step"] == steps
        close_rel(nump(st["exp_avg"]), z[f"{tag}_m_{i}"], f"{tag} exp_avg {i}")
        close_rel(nump(st["exp_avg_sq"]), z[f"{tag}_v_{i}"], f"{tag} exp_avg_sq {i}")
        close_rel(nump(st["slow_buffer"]), z[f"{tag}_slow_{i}"], f"{tag} slow_buffer {i}", rel=2e-6)


def test_ranger_weights_loaded_after_construction_follow_reference_trajectory():
    """the reference workflow: build the optimiser on the random-init net, THEN load pretrained weights without optimiser
    state (engine/train.py:52-55).  ranger2020.py:158-168 snapshots slow_buffer at a parameter's first step(), i.e. from
    the loaded weights -- so the golden trajectory (recorded from p0) must be reproduced when p0 arrives after Ranger()."""
    from tgpose_b200.ranger import Ranger
    z = golden("ranger")
    tag = "default"
    n, steps, snaps = int(z["n_tensors"]), int(z["steps"]), set(int(s) for s in z["snaps"])
    torch.manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(*z[f"p0_{i}"].shape).cuda()) for i in range(n)]       # "random init"
    opt = Ranger(params, **CONFIGS[tag])
    with torch.no_grad():
        for i, p in enumerate(params):                                                             # load_state_dict does copy_()
            p.copy_(torch.from_numpy(z[f"p0_{i}"]))
    for t in range(steps):
        opt.zero_grad()
        for i, p in enumerate(params):
            p.grad.copy_(torch.from_numpy(z[f"g_{t}_{i}"]))
        opt.clip_grad_norm_(5)
        opt.step()
        if t in snaps:                                     # the Lookahead pull-back at step k = 6 is inside the snapshots
            for i, p in enumerate(params):
                close_update(nump(p), z[f"{tag}_p_{t}_{i}"], z[f"p0_{i}"], f"late-load step {t + 1} tensor {i}")
    for i, p in enumerate(params):
        close_rel(nump(opt.state[p]["slow_buffer"]), z[f"{tag}_slow_{i}"], f"late-load slow_buffer {i}", rel=2e-6)
    with pytest.raises(RuntimeError):
        opt.add_param_group({"params": [torch.nn.Parameter(torch.zeros(3).cuda())]})


def test_ranger_groups_inactive_and_ragged_rows_vs_oracle():
    """two parameter groups with their own lr / weight decay, a tensor without a gradient (skipped like
    ranger2020.py:146-147), rows that are not 16-byte aligned, a 1-D tensor longer than one row piece, no clipping on
    odd steps -- against the numpy oracle run per group."""
    from tgpose_b200.ranger import Ranger
    g = torch.Generator().manual_seed(5)
    shapes_a = [(64, 1286, 1), (3, 7 * 128), (10001,), (17, 5, 3)]
    shapes_b = [(128, 8 * 64), (9,), (33, 2)]
    mk = lambda shapes: [torch.randn(*s, generator=g) * 0.2 for s in shapes]
    pa, pb, frozen = mk(shapes_a), mk(shapes_b), torch.randn(40, 3, generator=g)
    params_a = [torch.nn.Parameter(t.clone().cuda()) for t in pa]
    params_b = [torch.nn.Parameter(t.clone().cuda()) for t in pb]
    p_frozen = torch.nn.Parameter(frozen.clone().cuda())
    kw_a, kw_b = dict(lr=2e-3, weight_decay=0.0), dict(lr=5e-4, weight_decay=0.02)
    opt = Ranger([dict(params=params_a + [p_frozen], **kw_a), dict(params=params_b, **kw_b)], k=3)
    sa, sb = orc.ranger_init([t.numpy() for t in pa]), orc.ranger_init([t.numpy() for t in pb])
    for t in range(8):
        ga = [torch.randn(*s, generator=g) * (0.5 if t % 3 else 0.01) for s in shapes_a]
        gb = [torch.randn(*s, generator=g) * (0.5 if t % 3 else 0.01) for s in shapes_b]
        opt.zero_grad()
        for p, gr in zip(params_a + params_b, ga + gb):
            p.grad.copy_(gr)
        p_frozen.grad = None
        clip = 5.0 if t % 2 == 0 else None
        if clip:
            opt.clip_grad_norm_(clip)
        opt.step()
        allg = [x.numpy() for x in ga + gb]
        coef = np.float32(orc.clip_coef(allg, clip)[1]) if clip else np.float32(1.0)   # one norm over BOTH groups
        orc.ranger_step(sa, [(x.numpy() * coef).astype(np.float32) for x in ga], k=3, **kw_a)
        orc.ranger_step(sb, [(x.numpy() * coef).astype(np.float32) for x in gb], k=3, **kw_b)
    for i, p in enumerate(params_a):
        close_update(nump(p), sa["p"][i].astype(np.float64), pa[i].numpy(), f"group a tensor {i}")
        close_rel(nump(opt.state[p]["exp_avg_sq"]), sa["v"][i], f"group a exp_avg_sq {i}")
    for i, p in enumerate(params_b):
        close_update(nump(p), sb["p"][i].astype(np.float64), pb[i].numpy(), f"group b tensor {i}")
        close_rel(nump(opt.state[p]["exp_avg"]), sb["m"][i], f"group b exp_avg {i}")
    assert torch.equal(p_frozen.detach().cpu(), frozen)                                 # untouched
    assert opt.state[p_frozen]["step"] == 0
    assert float(opt.state[p_frozen]["exp_avg"].abs().max()) == 0.0


def test_ranger_picks_up_reallocated_gradients():
    """after a foreign zero_grad(set_to_none=True) autograd allocates fresh gradients outside the arena: step() copies
    them back in and re-points .grad at the arena."""
    from tgpose_b200.ranger import Ranger
    torch.manual_seed(1)
    lin = torch.nn.Linear(37, 11).cuda()
    w0 = lin.weight.detach().clone()
    opt = Ranger(lin.parameters(), lr=1e-2)
    lin.zero_grad(set_to_none=True)
    x = torch.randn(5, 37, device="cuda")
    lin(x).square().sum().backward()
    gw = lin.weight.grad.detach().clone()
    assert lin.weight.grad.data_ptr() != opt.flat_grads.data_ptr()
    opt.step()
    assert lin.weight.grad.data_ptr() == opt.flat_grads.data_ptr()
    st = orc.ranger_init([nump(w0)])
    orc.ranger_step(st, [nump(gw)], lr=1e-2)
    close_update(nump(lin.weight), st["p"][0].astype(np.float64), nump(w0), "weight after one step")


def test_ranger_rejects_cpu_parameters_and_bad_arguments():
    from tgpose_b200.ranger import Ranger
    with pytest.raises(RuntimeError):
        Ranger([torch.nn.Parameter(torch.zeros(3))])
    for bad in (dict(alpha=1.5), dict(k=0), dict(lr=0.0), dict(eps=0.0)):      # ranger2020.py:80-88
        with pytest.raises(ValueError):
            Ranger([torch.nn.Parameter(torch.zeros(3, device="cuda"))], **bad)


def test_ranger_state_dict_round_trip():
    """optimizer.state_dict() / load_state_dict() (trainer/RL_TDA.py:95,261): a resumed optimiser continues exactly
    like the original, with its state back inside the arenas."""
    from tgpose_b200.ranger import Ranger
    g = torch.Generator().manual_seed(11)
    shapes = [(8, 33), (33,), (4, 6, 1)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    grads = [[torch.randn(*s, generator=g) for s in shapes] for _ in range(8)]

    def run(opt, params, ts):
        for t in ts:
            opt.zero_grad()
            for p, gr in zip(params, grads[t]):
                p.grad.copy_(gr)
            opt.clip_grad_norm_(5)
            opt.step()

    pa = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    oa = Ranger(pa, lr=1e-2)
    run(oa, pa, range(5))
    sd = oa.state_dict()
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ob = Ranger(pb, lr=1e-2)
    ob.load_state_dict(sd)
    assert ob.steps == 5 and ob.state[pb[0]]["exp_avg"].data_ptr() == ob.flat_exp_avg.data_ptr()
    run(oa, pa, range(5, 8))
    run(ob, pb, range(5, 8))
    for a, b in zip(pa, pb):
        assert torch.equal(a.detach(), b.detach())
