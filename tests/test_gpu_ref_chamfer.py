"""GPU: tgp_chamfer_fwd / tgp_chamfer_bwd against the REFERENCE'S OWN KERNEL -- losses/chamfer3D/{chamfer_cuda.cpp,
chamfer3D.cu} compiled for sm_100a from where they lie by oracle/build_ref.py into oracle/_ref/chamfer3D (SURVEY 8c:
"GPU oracle and kernel to beat").  Distances and indices must be BIT-EQUAL on tie-free inputs: the arithmetic form
(difference, then fma(dz,dz, fma(dx,dx, dy*dy))) and the strict '<' of chamfer3D.cu:32-40 are part of the contract.
Skipped when oracle/_ref was not built (a checkout without /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    mod = build_ref.load_chamfer()
    if mod is None:
        pytest.skip("oracle/_ref/chamfer3D not built (no /root/reference at build time)")
    return mod


def _alloc(B, n, m, dev):
    return (torch.zeros(B, n, device=dev), torch.zeros(B, m, device=dev),
            torch.zeros(B, n, dtype=torch.int32, device=dev), torch.zeros(B, m, dtype=torch.int32, device=dev))


# (4,100,200): the reference's own unit-test shape (losses/metrics/CD/unit_test.py:14-35), all in the ragged tail path of
# chamfer3D.cu:88-131; 1028 x 1024: BASELINE configs[2]; 2500 x 3000: more than one 512-candidate batch + a ragged tail,
# more than one of our 2048-candidate tiles; (3,1,1), (2,5,70): degenerate sizes
@pytest.mark.parametrize("B,n,m", [(4, 100, 200), (8, 1028, 1024), (2, 2500, 3000), (3, 1, 1), (2, 5, 70), (1, 513, 2049)])
def test_forward_bit_equal_to_reference_kernel(ref, B, n, m):
    from tgpose_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + n + m)
    a = (torch.rand(B, n, 3, generator=g) - 0.5).to(dev)
    b = (torch.rand(B, m, 3, generator=g) - 0.5).to(dev)
    r, o = _alloc(B, n, m, dev), _alloc(B, n, m, dev)
    assert ref.forward(a, b, *r) == 1
    ops.chamfer_forward(a, b, *o)
    torch.cuda.synchronize()
    for name, x, y in zip(("dist1", "dist2", "idx1", "idx2"), r, o):
        assert torch.equal(x, y), f"{name}: {int((x != y).sum())} of {x.numel()} differ from the reference kernel"
    # and the CPU oracle (same arithmetic, oracle.c orc_chamfer_nn) agrees with both
    o1, o2, oi1, oi2 = orc.chamfer_forward(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(o1, o[0].cpu().numpy()) and np.array_equal(oi1, o[2].cpu().numpy())
    assert np.array_equal(o2, o[1].cpu().numpy()) and np.array_equal(oi2, o[3].cpu().numpy())


def test_forward_ties_take_the_lowest_index_like_the_reference(ref):
    """duplicated candidates: exact distance ties; strict '<' keeps the first (chamfer3D.cu:36)."""
    from tgpose_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    a = torch.rand(2, 300, 3, generator=g)
    b = torch.rand(2, 150, 3, generator=g)
    b = torch.cat([b, b, b[:, :40]], dim=1)           # every candidate appears 2-3 times
    a, b = a.to(dev), b.to(dev)
    r, o = _alloc(2, 300, 340, dev), _alloc(2, 300, 340, dev)
    ref.forward(a, b, *r)
    ops.chamfer_forward(a, b, *o)
    for x, y in zip(r, o):
        assert torch.equal(x, y)
    assert int(o[2].max()) < 150                      # always the first copy


def test_backward_matches_reference_kernel(ref):
    from tgpose_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    B, n, m = 4, 257, 300
    a, b = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, m, 3, generator=g).to(dev)
    gd1, gd2 = torch.randn(B, n, generator=g).to(dev), torch.randn(B, m, generator=g).to(dev)
    o = _alloc(B, n, m, dev)
    ops.chamfer_forward(a, b, *o)
    rg1, rg2 = torch.zeros(B, n, 3, device=dev), torch.zeros(B, m, 3, device=dev)
    og1, og2 = torch.zeros(B, n, 3, device=dev), torch.zeros(B, m, 3, device=dev)
    assert ref.backward(a, b, rg1, rg2, gd1, gd2, o[2], o[3]) == 1
    ops.chamfer_backward(a, b, gd1, gd2, o[2], o[3], og1, og2)
    torch.cuda.synchronize()
    # both accumulate the scattered terms with fp32 atomics (order not fixed): rel 1e-5 of the gradient scale
    for x, y in ((rg1, og1), (rg2, og2)):
        scale = float(x.abs().max())
        assert float((x - y).abs().max()) <= 1e-5 * scale
