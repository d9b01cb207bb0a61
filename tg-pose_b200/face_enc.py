"""Face_Enc -- the 3D-GCN backbone of TG-Pose (reference network/fs_net_repo/FaceRecon.py:12-86)
as one fused pipeline over the sm_100a kernels.

Same attribute / parameter names and construction order as the reference class (state_dicts and
seeds are interchangeable); the reference reads its hyper-parameters from absl FLAGS
(FaceRecon.py:15-17,54), here they are constructor arguments with the flags' defaults
(config/config.py:7,44-45,150).

What the fused forward does differently from calling the five layers one by one:
  * the xyz kNN is computed once per level and shared by conv RF-P / ORL / Pool (the reference
    recomputes identical results 3+1 times at level 0 and 2+1 times at level 1);
  * eval-mode BatchNorm + ReLU are folded into the last GEMM epilogue of each layer;
  * Pool gathers only the sampled rows.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gcn3d, ops


class Face_Enc(nn.Module):
    def __init__(self, neighbor_num=20, support_num=7, obj_c=6, output_channels=2500):
        super().__init__()
        self.neighbor_num = neighbor_num
        self.support_num = support_num
        self.output_channels = output_channels
        self.obj_c = obj_c

        self.conv_0 = gcn3d.HSlayer_surface(kernel_num=128, support_num=self.support_num)
        self.conv_1 = gcn3d.HS_layer(128, 128, support_num=self.support_num)
        self.pool_1 = gcn3d.Pool_layer(pooling_rate=4, neighbor_num=4)
        self.conv_2 = gcn3d.HS_layer(128, 256, support_num=self.support_num)
        self.conv_3 = gcn3d.HS_layer(256, 256, support_num=self.support_num)
        self.pool_2 = gcn3d.Pool_layer(pooling_rate=4, neighbor_num=4)
        self.conv_4 = gcn3d.HS_layer(256, 512, support_num=self.support_num)

        self.bn1 = nn.BatchNorm1d(128)
        self.bn2 = nn.BatchNorm1d(256)
        self.bn3 = nn.BatchNorm1d(256)

        feat_c = 128 + 128 + 256 + 256 + 512 + obj_c
        self.proj_layer = nn.Sequential(nn.Conv1d(feat_c, feat_c, kernel_size=1, bias=False),
                                        nn.BatchNorm1d(feat_c),
                                        nn.LeakyReLU(negative_slope=0.2),
                                        nn.Conv1d(feat_c, feat_c, kernel_size=1, bias=False))
        # test hooks (parity tiers T1/T2, SURVEY 8c'): replay / record the 12 kNN + 2 nearest index tensors
        self._inject = None
        self._record = None
        # CUDA-graph replay (graph.py): device tensors holding the two Pool_layer draws of this forward
        self._static_perms = None
        # inference: everything that depends on the coordinates only (level-1 / level-2 xyz kNN, both nearest-upsampling index
        # tensors) runs on a forked stream beside conv_0 / conv_1 instead of in the chain
        self.xyz_ahead = True
        self._xyz_stream = None

    # -- helpers -------------------------------------------------------------------------
    def _next_idx(self, compute):
        """index tensors in the reference's call order; each slot is either injected or computed."""
        if self._inject is not None:
            t = self._inject[self._slot].to(dtype=torch.int32).contiguous()
        else:
            t = compute()
        self._slot += 1
        if self._record is not None:
            self._record.append(t)
        return t

    def _bn_post(self, bn):
        """eval BatchNorm folded to per-channel (scale, shift) + ReLU for the GEMM epilogue; recomputed only when a
        parameter or running statistic of the layer changes."""
        def build():
            scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
            shift = bn.bias.detach() - bn.running_mean * scale
            return (scale.contiguous(), shift.contiguous(), True)
        return ops.PARAM_CACHE.get((bn.weight, bn.bias, bn.running_mean, bn.running_var), ("bn_post", bn.eps), build)

    def _bn_relu(self, bn, x):
        return F.relu(bn(x.transpose(1, 2)).transpose(1, 2))

    def _xyz_side_work(self, vertices, k):
        """The coordinate-only part of the forward on a forked stream.  Pool_layer's sample does not depend on the features
        (gcn3d.py:241-244: one torch.randperm per level), so the pooled clouds, their xyz kNN (conv_2 / conv_3 / conv_4 ORL,
        pool_2) and the two nearest-upsampling searches (FaceRecon.py:69-70) are known as soon as the input is.  Both draws are
        made here, in the reference's order.  Returns a dict; the caller waits for the stream before the first use."""
        dev = vertices.device
        n0 = vertices.shape[1]
        p1 = int(n0 / self.pool_1.pooling_rate)
        p2 = int(p1 / self.pool_2.pooling_rate)
        sp = self._static_perms
        if sp is not None:
            perm1, perm2 = sp
        else:
            perm1 = torch.randperm(n0)[:p1].to(dev, non_blocking=True)
            perm2 = torch.randperm(p1)[:p2].to(dev, non_blocking=True)
        if self._xyz_stream is None or self._xyz_stream.device != dev:
            self._xyz_stream = torch.cuda.Stream(device=dev)
        side, main = self._xyz_stream, torch.cuda.current_stream(dev)
        side.wait_stream(main)
        k1, k2 = min(k, p1 // 8), min(k, p2 // 8)
        with torch.cuda.stream(side):
            v1 = ops.select_rows(vertices, perm1)
            i_l1 = ops.knn_xyz(v1, k1, want64=False, want32=True)[1]
            v2 = ops.select_rows(v1, perm2)
            i_l2 = ops.knn_xyz(v2, k2, want64=False, want32=True)[1]
            nn1 = ops.nearest(vertices, v1, want64=False, want32=True)[1]
            nn2 = ops.nearest(vertices, v2, want64=False, want32=True)[1]
            up_rows = self.upsample_rows(nn1, nn2, p1, p2)
        return {"perm1": perm1, "perm2": perm2, "i_l1": i_l1, "i_l2": i_l2, "nn1": nn1, "nn2": nn2, "stream": side,
                "keep": (v1, v2), "up_rows": up_rows}

    @staticmethod
    def upsample_rows(nn1, nn2, n1, n2):
        """the rows of the level-1 / level-2 feature maps that every level-0 point upsamples from (FaceRecon.py:69-73) as GLOBAL
        row numbers (B*N,) int32 -- what the heads' factored first layers gather their coarse products by (posenet.py)."""
        B, N = nn1.shape[0], nn1.shape[1]
        cloud = torch.arange(B, device=nn1.device, dtype=torch.int32).view(B, 1)
        return ((nn1.reshape(B, N) + cloud * n1).reshape(-1).contiguous(),
                (nn2.reshape(B, N) + cloud * n2).reshape(-1).contiguous())

    # -- forward -------------------------------------------------------------------------
    def encode(self, vertices):
        """the five graph-conv feature maps and the two nearest-upsampling index tensors (FaceRecon.py:55-70):
        -> dict(fm_0 (B,N0,128), fm_1 (B,N0,128), fm_2, fm_3 (B,N1,256), fm_4 (B,N2,512), nn1, nn2 (B,N0,1) int32)."""
        k = self.neighbor_num
        # eval BatchNorm + ReLU folded into the GEMM epilogue only when nothing can ask for a gradient through it
        # (the folded constants are detached: an eval-mode forward with grad enabled takes the module path instead)
        fold = not self.training and not (torch.is_grad_enabled() and any(
            p.requires_grad for bn in (self.bn1, self.bn2, self.bn3) for p in bn.parameters()))
        self._slot = 0
        share = self._inject is None
        ahead = None
        if share and self.xyz_ahead and vertices.is_cuda and not torch.is_grad_enabled() and vertices.shape[1] >= 128:
            ahead = self._xyz_side_work(vertices.contiguous().float(), k)

        def xyz_knn(v, kk):
            return ops.knn_xyz(v, kk, want64=False, want32=True)[1]

        def feat_knn(f, kk, f_split=None):
            return ops.knn_feat(f, kk, want64=False, want32=True, x_split=f_split)[1]

        # level 0
        i0 = self._next_idx(lambda: xyz_knn(vertices, k))
        i0_orl = self._next_idx(lambda: i0 if share else xyz_knn(vertices, k))
        # HSlayer_surface uses one xyz index for both RF-P and ORL; with injected indices they are the same tensor
        fm_0, fm_0s = self.conv_0(vertices, k, idx_xyz=i0, post=(None, None, True), want_split=True)
        i1 = self._next_idx(lambda: feat_knn(fm_0, k, fm_0s))
        i1_orl = self._next_idx(lambda: i0 if share else xyz_knn(vertices, k))
        if fold:
            fm_1 = self.conv_1(vertices, fm_0, k, idx_feat=i1, idx_xyz=i1_orl, post=self._bn_post(self.bn1),
                               fm_split=fm_0s)
        else:
            fm_1 = self._bn_relu(self.bn1, self.conv_1(vertices, fm_0, k, idx_feat=i1, idx_xyz=i1_orl, fm_split=fm_0s))
        ip1 = self._next_idx(lambda: i0[:, :, :4].contiguous() if share else xyz_knn(vertices, 4))
        sp = self._static_perms
        if ahead is not None:
            sp = (ahead["perm1"], ahead["perm2"])
            torch.cuda.current_stream(vertices.device).wait_stream(ahead["stream"])
        v_pool_1, fm_pool_1 = self.pool_1(vertices, fm_1, idx_xyz=ip1, sample_idx=sp[0] if sp else None)

        # level 1
        k1 = min(k, v_pool_1.shape[1] // 8)
        i2 = self._next_idx(lambda: feat_knn(fm_pool_1, k1))
        i2_orl = self._next_idx(lambda: ahead["i_l1"] if ahead is not None else xyz_knn(v_pool_1, k1))
        if fold:
            fm_2, fm_2s = self.conv_2(v_pool_1, fm_pool_1, k1, idx_feat=i2, idx_xyz=i2_orl,
                                      post=self._bn_post(self.bn2), want_split=True)
        else:
            fm_2s = None
            fm_2 = self._bn_relu(self.bn2, self.conv_2(v_pool_1, fm_pool_1, k1, idx_feat=i2, idx_xyz=i2_orl))
        i3 = self._next_idx(lambda: feat_knn(fm_2, k1, fm_2s))
        i3_orl = self._next_idx(lambda: i2_orl if share else xyz_knn(v_pool_1, k1))
        if fold:
            fm_3 = self.conv_3(v_pool_1, fm_2, k1, idx_feat=i3, idx_xyz=i3_orl, post=self._bn_post(self.bn3),
                               fm_split=fm_2s)
        else:
            fm_3 = self._bn_relu(self.bn3, self.conv_3(v_pool_1, fm_2, k1, idx_feat=i3, idx_xyz=i3_orl))
        ip2 = self._next_idx(lambda: (i2_orl[:, :, :4].contiguous() if (share and k1 >= 4) else xyz_knn(v_pool_1, 4)))
        v_pool_2, fm_pool_2 = self.pool_2(v_pool_1, fm_3, idx_xyz=ip2, sample_idx=sp[1] if sp else None)

        # level 2
        k2 = min(k, v_pool_2.shape[1] // 8)
        i4 = self._next_idx(lambda: feat_knn(fm_pool_2, k2))
        i4_orl = self._next_idx(lambda: ahead["i_l2"] if ahead is not None else xyz_knn(v_pool_2, k2))
        fm_4 = self.conv_4(v_pool_2, fm_pool_2, k2, idx_feat=i4, idx_xyz=i4_orl)

        # nearest-neighbour upsampling indices back to level 0 (FaceRecon.py:69-70)
        nn1 = self._next_idx(lambda: ahead["nn1"] if ahead is not None else ops.nearest(vertices, v_pool_1, want64=False, want32=True)[1])
        nn2 = self._next_idx(lambda: ahead["nn2"] if ahead is not None else ops.nearest(vertices, v_pool_2, want64=False, want32=True)[1])
        return {"fm_0": fm_0, "fm_1": fm_1, "fm_2": fm_2, "fm_3": fm_3, "fm_4": fm_4, "nn1": nn1, "nn2": nn2,
                "up_rows": ahead["up_rows"] if ahead is not None else None}

    def one_hot(self, cat_id, bs):
        obj_idh = cat_id.view(-1, 1)
        return torch.zeros(bs, self.obj_c, device=cat_id.device).scatter_(1, obj_idh.long(), 1)

    @staticmethod
    def concat_sources(parts, one_hot, extra=()):
        """the column blocks of `feat` (FaceRecon.py:81) as tgp_concat_rows sources (+ optional extra blocks)."""
        f = parts
        n1, n2 = f["fm_2"].shape[1], f["fm_4"].shape[1]
        flat = lambda t: t.reshape(-1, t.shape[-1])
        nn1, nn2 = f["nn1"].reshape(f["nn1"].shape[0], -1), f["nn2"].reshape(f["nn2"].shape[0], -1)
        return [(flat(f["fm_0"]), None, 1), (flat(f["fm_1"]), None, 1), (flat(f["fm_2"]), nn1, n1),
                (flat(f["fm_3"]), nn1, n1), (flat(f["fm_4"]), nn2, n2), (one_hot, None, 0)] + list(extra)

    def forward(self, vertices, cat_id, enable_proj=False):
        """vertices (B,N,3), cat_id (B,1) -> (feat (B,N,1286), feat_global (B,1286,N)); ref FaceRecon.py:39-86."""
        bs, vertice_num, _ = vertices.size()
        parts = self.encode(vertices)
        one_hot = self.one_hot(cat_id, bs)
        if torch.is_grad_enabled() and any(parts[k].requires_grad for k in ("fm_0", "fm_1", "fm_2", "fm_3", "fm_4")):
            # autograd path: the gathers scatter-add in backward (GatherRowsFn), the concatenation is torch's
            fm_2u = gcn3d.indexing_neighbor_new(parts["fm_2"], parts["nn1"]).squeeze(2)
            fm_3u = gcn3d.indexing_neighbor_new(parts["fm_3"], parts["nn1"]).squeeze(2)
            fm_4u = gcn3d.indexing_neighbor_new(parts["fm_4"], parts["nn2"]).squeeze(2)
            feat = torch.cat([parts["fm_0"], parts["fm_1"], fm_2u, fm_3u, fm_4u,
                              one_hot.unsqueeze(1).expand(-1, vertice_num, -1)], dim=2)
        else:
            # one launch: upsampling gathers + one-hot broadcast + concatenation
            feat = ops.concat_rows(self.concat_sources(parts, one_hot), bs, vertice_num)[0].view(bs, vertice_num, -1)
        feat_global = feat.permute(0, 2, 1)
        feat_global_prj = self.proj_layer(feat_global) if enable_proj else feat_global
        return feat, feat_global_prj
