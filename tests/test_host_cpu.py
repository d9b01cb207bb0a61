"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/tgpose_b200.h
declares; the host-side mirrors keep the reference's parameter names / shapes; no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tgpose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tgpose_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tgpose_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.tgp_version() >= 100
    assert isinstance(lib.tgp_launch_count(), int)


def test_gemm_args_struct_layout_matches_header():
    from tgpose_b200 import _lib
    # natural alignment of the C structs (LP64): tgp_out_seg 32 B, tgp_gemm_args 152 + 4*32 + 16 + (int mixed, a_kp,
    # a_group_cols, padded to 16)
    assert ctypes.sizeof(_lib.OutSeg) == 32
    assert ctypes.sizeof(_lib.GemmArgs) == 152 + 4 * 32 + 16 + 16
    assert _lib.GemmArgs.seg.offset == 152
    assert _lib.GemmArgs.mixed.offset == 152 + 4 * 32 + 16
    assert _lib.GemmArgs.a_kp.offset == _lib.GemmArgs.mixed.offset + 4
    assert _lib.GemmArgs.a_group_cols.offset == _lib.GemmArgs.mixed.offset + 8
    assert ctypes.sizeof(_lib.ConcatSrc) == 40


def test_no_cpu_fallback():
    from tgpose_b200 import gcn3d
    x = torch.rand(1, 16, 3)
    with pytest.raises(RuntimeError):
        gcn3d.get_neighbor_index(x, 4)
    layer = gcn3d.HS_layer(8, 8, 7)
    with pytest.raises(RuntimeError):
        layer(x, torch.rand(1, 16, 8), 4)


def test_state_dict_names_and_shapes():
    """parameter names/shapes are API (checkpoints, trainer/RL_TDA.py:64-97)."""
    from tgpose_b200.face_enc import Face_Enc
    torch.manual_seed(0)
    sd = Face_Enc().state_dict()
    assert tuple(sd["conv_0.directions"].shape) == (3, 896)
    assert tuple(sd["conv_0.STE_layer.weight"].shape) == (128, 3, 1)
    assert tuple(sd["conv_0.conv2.weight"].shape) == (128, 256, 1)
    for name, (cin, cout) in {"conv_1": (128, 128), "conv_2": (128, 256), "conv_3": (256, 256), "conv_4": (256, 512)}.items():
        assert tuple(sd[f"{name}.weights"].shape) == (cin, 8 * cout)
        assert tuple(sd[f"{name}.bias"].shape) == (8 * cout,)
        assert tuple(sd[f"{name}.directions"].shape) == (3, 7 * cout)
        assert tuple(sd[f"{name}.STE_layer.weight"].shape) == (cout, cin, 1)
        assert tuple(sd[f"{name}.conv2.weight"].shape) == (cout, 2 * cout, 1)
    assert "bn1.running_mean" in sd and "proj_layer.0.weight" in sd


def test_pack_layer_column_order():
    """the slab column permutation (cgroup, s, c4) used by the projection GEMM."""
    from tgpose_b200.autograd import _pack_layer
    S, C, cin = 7, 8, 5
    w = torch.arange(cin * (S + 1) * C, dtype=torch.float32).reshape(cin, (S + 1) * C)
    b = torch.arange((S + 1) * C, dtype=torch.float32)
    ste = torch.arange(C * cin, dtype=torch.float32).reshape(C, cin, 1) + 1000
    wcat, bcat, _ = _pack_layer(w, b, ste, S, C)
    assert wcat.shape == (cin, (S + 2) * C) and bcat.shape == ((S + 2) * C,)
    assert torch.equal(wcat[:, :C], w[:, :C])
    for cg in range(C // 4):
        for s in range(S):
            for c4 in range(4):
                col = C + cg * S * 4 + s * 4 + c4
                assert torch.equal(wcat[:, col], w[:, C + s * C + cg * 4 + c4])
                assert bcat[col] == b[C + s * C + cg * 4 + c4]
    assert torch.equal(wcat[:, (S + 1) * C:], ste[:, :, 0].t())
    assert (bcat[(S + 1) * C:] == 0).all()


def test_pool_consumes_cpu_rng_like_reference():
    """Pool_layer draws torch.randperm(N) from the global CPU generator (gcn3d.py:242)."""
    torch.manual_seed(7)
    a = torch.randperm(128)[:32]
    torch.manual_seed(7)
    b = torch.randperm(128)
    c = torch.randperm(32)
    torch.manual_seed(7)
    assert torch.equal(torch.randperm(128)[:32], a)
    assert torch.equal(torch.randperm(32), c)
    g = np.load(os.path.join(ROOT, "tests", "golden", "convs.npz"))
    assert np.array_equal(b.numpy().astype(np.int16), g["p_perm"])
