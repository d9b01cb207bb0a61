"""Drop-in for the reference's network/fs_net_repo/gcn3d.py on sm_100a kernels.

Same public names, positional signatures, tensor layouts and parameter names/shapes
(`directions`, `weights`, `bias`, `STE_layer.weight`, `conv2.weight` -- checkpoints load
unchanged), same constructor RNG consumption (the same torch.manual_seed yields
the same initial weights), same CPU-RNG `torch.randperm` draw in Pool_layer (gcn3d.py:242).
`FaceRecon.py` runs on it by swapping one import:
    import tgpose_b200.gcn3d as gcn3d      # was: import network.fs_net_repo.gcn3d as gcn3d

What differs is where the work happens: every function below is a handful of launches into
libtgpose_b200.so; nothing of size (B,N,N) or (B,N,k,S*C) is materialised.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .autograd import GatherRowsFn, HSLayerFn, HSSurfaceFn, PoolFn


# ----------------------------------------------------------------------------- free functions
def get_neighbor_index(vertices: "(bs, vertice_num, dim)", neighbor_num: int):
    """ref gcn3d.py:14-23 -> (bs, vertice_num, neighbor_num) int64."""
    if vertices.shape[-1] == 3:
        return ops.knn_xyz(vertices, neighbor_num)[0]
    return ops.knn_feat(vertices, neighbor_num)[0]


def get_nearest_index(target: "(bs, v1, 3)", source: "(bs, v2, 3)"):
    """ref gcn3d.py:26-35 -> (bs, v1, 1) int64."""
    return ops.nearest(target, source)[0]


def indexing_neighbor_new(tensor: "(bs, vertice_num, dim)", index: "(bs, vertice_num, neighbor_num)"):
    """ref gcn3d.py:38-46 -> (bs, vertice_num, neighbor_num, dim)."""
    if torch.is_grad_enabled() and tensor.requires_grad:
        return GatherRowsFn.apply(tensor, index)
    return ops.gather_rows(tensor, index)


def get_neighbor_direction_norm(vertices, neighbor_index, return_unnormed=False):
    """ref gcn3d.py:48-58 -> (bs, vertice_num, neighbor_num, 3)."""
    norm = ops.direction_norm(vertices, neighbor_index)
    if return_unnormed:
        return norm, ops.gather_rows(vertices, neighbor_index) - vertices.unsqueeze(2)
    return norm


def get_receptive_fields(neighbor_num, vertices, feature_map=None, mode='RF-F'):
    """ref gcn3d.py:188-208 -> (direction_norm, neighbor_index)."""
    assert mode in ['RF-F', 'RF-P']
    if mode == 'RF-F':
        assert feature_map is not None, "The feature_map should be provided if 'RF-F' is used"
        feat = feature_map
    else:
        feat = vertices
    neighbor_index = get_neighbor_index(feat, neighbor_num)
    return get_neighbor_direction_norm(vertices, neighbor_index), neighbor_index


def get_ORL_global(feature, vertices, neighbor_num):
    """ref gcn3d.py:210-217 -> (bs, vertice_num, channel): the cloud-global feature repeated per point."""
    idx32 = ops.knn_xyz(vertices, neighbor_num, want64=False, want32=True)[1]
    g = ops.orl_global(feature, idx32)
    return g.unsqueeze(1).repeat(1, feature.size(1), 1)


# ----------------------------------------------------------------------------- modules
class HSlayer_surface(nn.Module):
    """ref gcn3d.py:60-112 (upstream 3D-GCN "Conv_surface")."""

    def __init__(self, kernel_num, support_num):
        super().__init__()
        self.feat_k = 8
        self.kernel_num = kernel_num
        self.support_num = support_num
        self.relu = nn.ReLU(inplace=True)
        self.directions = nn.Parameter(torch.empty(3, support_num * kernel_num))
        self.STE_layer = nn.Conv1d(3, kernel_num, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(2 * kernel_num, kernel_num, kernel_size=1, bias=False)
        self.initialize()

    def initialize(self):
        stdv = 1. / math.sqrt(self.support_num * self.kernel_num)
        self.directions.data.uniform_(-stdv, stdv)

    def forward(self, vertices: "(bs, vertice_num, 3)", neighbor_num: int, idx_xyz=None, post=None, want_split=False):
        """-> (bs, vertice_num, kernel_num).  idx_xyz / post / want_split are extensions used by the fused
        encoder: a precomputed int32 xyz kNN, (scale, shift, relu) folded into the last epilogue, and the
        result also returned as the next projection's tensor-core operand (-> (out, out_split))."""
        out, out_split = HSSurfaceFn.apply(vertices, self.directions, self.STE_layer.weight, self.conv2.weight,
                                           neighbor_num, self.support_num, self.kernel_num, idx_xyz, post, want_split,
                                           torch.is_grad_enabled())
        return (out, out_split) if want_split else out

    def graph_conv(self, receptive_fields_norm, vertices, neighbor_num):
        """ref gcn3d.py:91-106 (kept for API parity; recomputes the xyz kNN like get_receptive_fields)."""
        idx32 = ops.knn_xyz(vertices, neighbor_num, want64=False, want32=True)[1]
        return ops.surface_conv(vertices, idx32, self.directions, self.support_num, self.kernel_num)


class HS_layer(nn.Module):
    """ref gcn3d.py:115-186 (upstream 3D-GCN "Conv_layer")."""

    def __init__(self, in_channel, out_channel, support_num):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.support_num = support_num
        self.relu = nn.ReLU(inplace=True)
        self.weights = nn.Parameter(torch.empty(in_channel, (support_num + 1) * out_channel))
        self.bias = nn.Parameter(torch.empty((support_num + 1) * out_channel))
        self.directions = nn.Parameter(torch.empty(3, support_num * out_channel))
        self.feat_k = 8
        self.STE_layer = nn.Conv1d(self.in_channel, self.out_channel, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(2 * out_channel, out_channel, kernel_size=1, bias=False)
        self.initialize()

    def initialize(self):
        stdv = 1. / math.sqrt(self.out_channel * (self.support_num + 1))
        self.weights.data.uniform_(-stdv, stdv)
        self.bias.data.uniform_(-stdv, stdv)
        self.directions.data.uniform_(-stdv, stdv)

    def forward(self, vertices: "(bs, vertice_num, 3)", feature_map: "(bs, vertice_num, in_channel)",
                neighbor_num: int, idx_feat=None, idx_xyz=None, post=None, fm_split=None, want_split=False):
        """-> (bs, vertice_num, out_channel).  idx_feat / idx_xyz (int32), post, fm_split (feature_map already
        split for the tensor cores) and want_split (-> (out, out_split)) are fused-encoder extensions."""
        out, out_split = HSLayerFn.apply(vertices, feature_map, self.weights, self.bias, self.directions,
                                         self.STE_layer.weight, self.conv2.weight, neighbor_num, self.support_num,
                                         self.out_channel, idx_feat, idx_xyz, post, fm_split, want_split,
                                         torch.is_grad_enabled())
        return (out, out_split) if want_split else out


class Pool_layer(nn.Module):
    """ref gcn3d.py:219-245."""

    def __init__(self, pooling_rate: int = 4, neighbor_num: int = 4):
        super().__init__()
        self.pooling_rate = pooling_rate
        self.neighbor_num = neighbor_num

    def forward(self, vertices: "(bs, vertice_num, 3)", feature_map: "(bs, vertice_num, channel_num)", idx_xyz=None,
                sample_idx=None):
        """-> (vertices_pool (bs, n/rate, 3), feature_map_pool (bs, n/rate, channel)).
        The sample is one torch.randperm on the global CPU generator shared by the batch (gcn3d.py:241-244).
        sample_idx (extension): a device tensor that already holds that draw (CUDA-graph replay, graph.py)."""
        vertice_num = vertices.size(1)
        pool_num = int(vertice_num / self.pooling_rate)
        if sample_idx is None:
            sample_idx = torch.randperm(vertice_num)[:pool_num]
        return PoolFn.apply(vertices, feature_map, sample_idx, self.neighbor_num, idx_xyz, torch.is_grad_enabled())
