"""PoseNet9D -- the full TG-Pose network (reference network/fs_net_repo/PoseNet9D.py:19-91) with the
3D-GCN backbone on the sm_100a kernels.

Only the backbone (Face_Enc) is the hot path of this build.  The heads -- Face_Dec / PH_Predictor
(FaceRecon.py:89-167), Rot_green / Rot_red (PoseR.py:10-69), Pose_Ts (PoseTs.py:13-45) -- are the
reference's plain Conv1d/BatchNorm/Linear stacks; they are restated here (same attribute names,
same construction order, so state_dicts and seeds are interchangeable) because the reference
package is not importable on the GPU box.  Their 1x1 convolutions run through `pointwise()`:
the library's GEMM with the eval-BatchNorm + ReLU epilogue in inference, torch layers in training.

Hyper-parameters that the reference reads from absl FLAGS are constructor arguments with the
flags' defaults (config/config.py:7,32-38,44-45,150).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .face_enc import Face_Enc

FEAT_C = 1286  # 128 + 128 + 256 + 256 + 512 + obj_c(6), FLAGS.feat_c_R


def _fold(conv, bn):
    """(scale, shift) of conv-bias + eval BatchNorm as one per-channel affine applied to W.x."""
    cout = conv.out_channels
    dev = conv.weight.device
    if bn is None:
        scale = torch.ones(cout, device=dev)
        shift = conv.bias.detach().clone() if conv.bias is not None else torch.zeros(cout, device=dev)
        return scale, shift
    scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() - bn.running_mean * scale
    if conv.bias is not None:
        shift = shift + conv.bias.detach() * scale
    return scale, shift



def _require_cuda(t, what):
    """the heads run on the library kernels only: there is no CPU / eager fallback (north_star)."""
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required -- tg-pose_b200 has no CPU path")

class _Packed:
    """weights of one fused 1x1-conv stage: (Ncols, K) matrix, its tensor-core split, affine and activation."""

    def __init__(self, mats, folds, slopes, k_pad_to=None):
        ws = []
        for w in mats:
            w = w.detach().reshape(w.shape[0], w.shape[1])
            if k_pad_to is not None and w.shape[1] < k_pad_to:
                w = torch.cat([w, w.new_zeros(w.shape[0], k_pad_to - w.shape[1])], 1)
            ws.append(w)
        self.w = torch.cat(ws, 0).contiguous()
        self.scale = torch.cat([f[0] for f in folds]).contiguous()
        self.shift = torch.cat([f[1] for f in folds]).contiguous()
        self.slope = torch.cat([torch.full((m.shape[0],), float(sl), device=self.w.device)
                                for m, sl in zip(mats, slopes)]).contiguous()
        # the heads' contractions run on MIXED operands (TF32 + bf16 cross terms): two TF32-pass equivalents instead of
        # three at ~2^-19 relative error per product; nothing downstream of the heads is a neighbour search
        # (weight operand of a mixed = 2 contraction: residual slot in fp16, so the activations need no bf16(x) slot)
        self.w_split = ops.split_mixed(self.w, w16=True)

    @classmethod
    def from_matrix(cls, w, scale=None, shift=None, slope=None):
        """a pack from an already assembled (Ncols, K) matrix (column subsets of a layer's weight)."""
        pk = cls.__new__(cls)
        pk.w = w.detach().contiguous()
        pk.scale = scale.contiguous() if scale is not None else None
        pk.shift = shift.contiguous() if shift is not None else None
        pk.slope = slope.contiguous() if slope is not None else None
        pk.w_split = ops.split_mixed(pk.w, w16=True)
        return pk


# column blocks of `feat` (FaceRecon.py:81) [fm_0 | fm_1 | up(fm_2) | up(fm_3) | up(fm_4) | one-hot] (+ xyz for Pose_Ts)
_FINE_COLS = list(range(0, 256)) + list(range(1280, FEAT_C + 3))      # per-point blocks: fm_0, fm_1, one-hot, xyz
_L1_COLS = slice(256, 768)                                             # fm_2 | fm_3: level-1 points, upsampled by nn1
_L2_COLS = slice(768, 1280)                                            # fm_4: level-2 points, upsampled by nn2


def pointwise(x_cl, conv, bn=None, act=None, slope=0.2):
    """1x1 Conv1d (+ eval BatchNorm + activation) on a channel-last (B,N,Cin) tensor -> (B,N,Cout): one GEMM
    launch with the affine and the activation folded into the epilogue."""
    B, N, cin = x_cl.shape
    scale, shift = _fold(conv, bn)
    ns = None
    if act == "leaky":
        ns = torch.full((conv.out_channels,), slope, device=x_cl.device)
    out = ops.linear_nk(x_cl.reshape(B * N, cin), conv.weight.reshape(conv.out_channels, cin), scale=scale,
                        shift=shift, relu=(act == "relu"), neg_slope=ns)
    return out.view(B, N, conv.out_channels)


class Face_Dec(nn.Module):
    """ref FaceRecon.py:89-117."""

    def __init__(self, dim_fuse):
        super().__init__()
        self.recon_num = 3
        self.conv1d_block = nn.Sequential(
            nn.Conv1d(dim_fuse, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
            nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
            nn.Conv1d(512, 256, 1), nn.BatchNorm1d(256), nn.ReLU(inplace=True))
        self.recon_head = nn.Sequential(
            nn.Conv1d(256, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
            nn.Conv1d(128, self.recon_num, 1))

    def forward(self, x):
        """x (B, C, N) -> recon (B, N, 3)."""
        _require_cuda(x, "FaceRecon decoder")
        if self.training:
            return self.forward_rows(x.permute(0, 2, 1).contiguous())
        h = x.permute(0, 2, 1).contiguous()
        b = self.conv1d_block
        h = pointwise(h, b[0], b[1], "relu")
        h = pointwise(h, b[3], b[4], "relu")
        h = pointwise(h, b[6], b[7], "relu")
        r = self.recon_head
        h = pointwise(h, r[0], r[1], "relu")
        return pointwise(h, r[3])

    def forward_rows(self, x_cl, x_split=None):
        """train mode on channel-last input (B, N, C) -> recon (B, N, 3): the four Conv1d-BN-ReLU blocks on the library's
        kernels (heads_train.py), the final 128 -> 3 projection as a torch linear."""
        B, N, C = x_cl.shape
        b, r = self.conv1d_block, self.recon_head
        h = _train_chain(x_cl.reshape(B * N, C), [(b[0], b[1], 0.0), (b[3], b[4], 0.0), (b[6], b[7], 0.0), (r[0], r[1], 0.0)],
                         x_split)
        last = r[3]
        return F.linear(h, last.weight.reshape(last.out_channels, -1), last.bias).view(B, N, -1)


def _train_chain(x2d, layers, x_split=None):
    """Conv1d-BN-ReLU blocks in train mode on channel-last rows; every block hands the next one its split operand."""
    from .heads_train import conv_bn_act_train
    h, hs = x2d, x_split
    for i, (conv, bn, slope) in enumerate(layers):
        last = i == len(layers) - 1
        if last:
            h = conv_bn_act_train(h, conv, bn, slope, x_split=hs)
        else:
            h, hs = conv_bn_act_train(h, conv, bn, slope, x_split=hs, want_split=True)
    return h


class PH_Predictor(nn.Module):
    """ref FaceRecon.py:120-167."""

    def __init__(self, output_channels=2500):
        super().__init__()
        self.output_channels = output_channels
        self.conv_5 = nn.Sequential(nn.Conv1d(FEAT_C, 1024, kernel_size=1, bias=False),
                                    nn.BatchNorm1d(1024),
                                    nn.LeakyReLU(negative_slope=0.2))
        self.linear1 = nn.Linear(1024 * 2, 1024, bias=False)
        self.bn5 = nn.BatchNorm1d(1024)
        self.dp1 = nn.Dropout(p=0.5)
        self.linear2 = nn.Linear(1024, self.output_channels)
        self.linear3 = nn.Linear(1024, self.output_channels)
        self.linear4 = nn.Linear(self.output_channels, FEAT_C)
        self.linear5 = nn.Linear(self.output_channels, FEAT_C)
        self.ac2 = nn.Sigmoid()
        self.ac3 = nn.Sigmoid()

    def forward(self, feat):
        """feat (B,N,1286) -> (feat + pi1 + pi2 as (B,1286,N), h1, h2)."""
        bs = feat.shape[0]
        _require_cuda(feat, "PH_Predictor")
        if self.training:
            f = _train_chain(feat.reshape(-1, feat.shape[2]), [(self.conv_5[0], self.conv_5[1], 0.2)])
            pooled = f.view(bs, feat.shape[1], -1).max(dim=1)[0]
        else:
            f = pointwise(feat, self.conv_5[0], self.conv_5[1], "leaky")
            pooled = f.max(dim=1)[0]
        feat_all = torch.cat((pooled, pooled), 1)      # the reference pools the same tensor twice (FaceRecon.py:146-148)
        feat_all = F.leaky_relu(self.bn5(self.linear1(feat_all)), negative_slope=0.2)
        feat_all = self.dp1(feat_all)
        pi1 = self.linear2(feat_all)
        pi1_1 = self.linear4(pi1)
        h1 = self.ac2(pi1)
        pi2 = self.linear3(feat_all)
        pi2_1 = self.linear5(pi2)
        h2 = self.ac3(pi2)
        out = feat.permute(0, 2, 1) + (pi1_1 + pi2_1).unsqueeze(-1)
        return out, h1, h2


class FaceNet(nn.Module):
    """ref FaceRecon.py:170-202."""

    def __init__(self, **enc_kwargs):
        super().__init__()
        self.encoder = Face_Enc(**enc_kwargs)
        self.decoder = Face_Dec(FEAT_C)
        self.ph_pred = PH_Predictor(enc_kwargs.get("output_channels", 2500))

    def forward(self, vertices, cat_id, enable_proj=False, pred_PH=True):
        feat, feat_global = self.encoder(vertices, cat_id, enable_proj)
        if pred_PH:
            feat_ph, h1, h2 = self.ph_pred(feat)
            recon = self.decoder(feat_ph)
        else:
            recon = self.decoder(feat.permute(0, 2, 1))
            h1, h2 = None, None
        return recon, feat, feat_global, h1, h2


class _PointHead(nn.Module):
    """Conv1d(f,1024)-BN-ReLU, Conv1d(1024,256)-BN-ReLU, max over points, Conv1d(256,256)-BN-ReLU,
    Dropout, Conv1d(256,k): the shared shape of Rot_green / Rot_red (PoseR.py:10-69) and Pose_Ts (PoseTs.py:13-45)."""

    def __init__(self, f, k):
        super().__init__()
        self.f = f
        self.k = k
        self.conv1 = nn.Conv1d(self.f, 1024, 1)
        self.conv2 = nn.Conv1d(1024, 256, 1)
        self.conv3 = nn.Conv1d(256, 256, 1)
        self.conv4 = nn.Conv1d(256, self.k, 1)
        self.drop1 = nn.Dropout(0.2)
        self.bn1 = nn.BatchNorm1d(1024)
        self.bn2 = nn.BatchNorm1d(256)
        self.bn3 = nn.BatchNorm1d(256)

    def trunk(self, x):
        """x (B, f, N) -> (B, k)."""
        _require_cuda(x, "pose head trunk")
        if self.training:
            B, C, N = x.shape
            h = _train_chain(x.permute(0, 2, 1).reshape(B * N, C), [(self.conv1, self.bn1, 0.0), (self.conv2, self.bn2, 0.0)])
            x = h.view(B, N, -1).max(dim=1)[0].unsqueeze(2)
        else:
            h = pointwise(x.permute(0, 2, 1).contiguous(), self.conv1, self.bn1, "relu")
            h = pointwise(h, self.conv2, self.bn2, "relu")
            x = h.max(dim=1)[0].unsqueeze(2)
        x = F.relu(self.bn3(self.conv3(x)))
        x = self.drop1(x)
        x = self.conv4(x)
        return x.squeeze(2).contiguous()


class Rot_green(_PointHead):
    def __init__(self, feat_c_R=FEAT_C, R_c=4):
        super().__init__(feat_c_R, R_c)

    def forward(self, x):
        return self.trunk(x)


class Rot_red(_PointHead):
    def __init__(self, feat_c_R=FEAT_C, R_c=4):
        super().__init__(feat_c_R, R_c)

    def forward(self, x):
        return self.trunk(x)


class Pose_Ts(_PointHead):
    def __init__(self, feat_c_ts=FEAT_C + 3, Ts_c=6):
        super().__init__(feat_c_ts, Ts_c)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()
        self.relu3 = nn.ReLU()

    def forward(self, x):
        x = self.trunk(x)
        return x[:, 0:3], x[:, 3:6]


class PoseNet9D(nn.Module):
    """ref PoseNet9D.py:19-91.  `train_outputs` plays the role of FLAGS.train (PoseNet9D.py:68)."""

    def __init__(self, only_encoder=False, train_outputs=False, **enc_kwargs):
        super().__init__()
        self.only_encoder = only_encoder
        self.train_outputs = train_outputs
        self.overlap_heads = True      # inference: pose tails on forked streams next to the decoder chain
        self.factored_heads = True     # inference: upsampled input channels of the first layers contracted per COARSE point
        self._side = None
        if not only_encoder:
            self.face_all = FaceNet(**enc_kwargs)
            self.rot_green = Rot_green()
            self.rot_red = Rot_red()
            self.ts = Pose_Ts()
        else:
            self.face_enc = FaceNet(**enc_kwargs)

    # ---- inference: every per-point 1x1 convolution of the heads as fused tensor-core GEMMs ------------
    def _head_packs(self):
        """packed / BN-folded / tf32-split head weights, rebuilt only when a parameter or BN buffer changes."""
        fa, g, r, t = self.face_all, self.rot_green, self.rot_red, self.ts
        d, ph = fa.decoder, fa.ph_pred
        mods = [g, r, t, d, ph]
        key = tuple((x._version, x.data_ptr()) for m in mods for x in list(m.parameters()) + list(m.buffers()))
        if getattr(self, "_packs_key", None) != key:
            kin = FEAT_C + 3
            c = d.conv1d_block
            dev = g.conv1.weight.device
            with torch.no_grad():
                ones = lambda n: torch.ones(n, device=dev)
                zeros = lambda n: torch.zeros(n, device=dev)
                w_l1 = ph.linear1.weight.detach()
                bn5_scale = ph.bn5.weight.detach() * torch.rsqrt(ph.bn5.running_var + ph.bn5.eps)
                self._packs = {
                    # [rot_green.conv1 | rot_red.conv1 | ph_pred.conv_5 | ts.conv1] on [feat | xyz] (K = 1289)
                    "stage1": _Packed([g.conv1.weight, r.conv1.weight, ph.conv_5[0].weight, t.conv1.weight],
                                      [_fold(g.conv1, g.bn1), _fold(r.conv1, r.bn1), _fold(ph.conv_5[0], ph.conv_5[1]),
                                       _fold(t.conv1, t.bn1)], [0.0, 0.0, 0.2, 0.0], k_pad_to=kin),
                    "green2": _Packed([g.conv2.weight], [_fold(g.conv2, g.bn2)], [0.0]),
                    "red2": _Packed([r.conv2.weight], [_fold(r.conv2, r.bn2)], [0.0]),
                    "ts2": _Packed([t.conv2.weight], [_fold(t.conv2, t.bn2)], [0.0]),
                    "dec1": _Packed([c[0].weight], [_fold(c[0], c[1])], [0.0], k_pad_to=kin),
                    "dec2": _Packed([c[3].weight], [_fold(c[3], c[4])], [0.0]),
                    "dec3": _Packed([c[6].weight], [_fold(c[6], c[7])], [0.0]),
                    "dec4": _Packed([d.recon_head[0].weight], [_fold(d.recon_head[0], d.recon_head[1])], [0.0]),
                    # per-cloud tails (M = batch): plain fp32 matrices for the skinny kernel
                    # linear1 sees cat(pooled, pooled) (FaceRecon.py:146-149) = pooled @ (W[:, :1024] + W[:, 1024:])^T
                    "ph_l1": ((w_l1[:, :1024] + w_l1[:, 1024:]).contiguous(), bn5_scale.contiguous(),
                              (ph.bn5.bias.detach() - ph.bn5.running_mean * bn5_scale).contiguous(),
                              torch.full((1024,), 0.2, device=dev)),
                    "ph_l23": (torch.cat([ph.linear2.weight.detach(), ph.linear3.weight.detach()], 0).contiguous(),
                               torch.cat([ph.linear2.bias.detach(), ph.linear3.bias.detach()]).contiguous()),
                    "ph_l45": (torch.cat([ph.linear4.weight.detach(), ph.linear5.weight.detach()], 1).contiguous(),
                               (ph.linear4.bias.detach() + ph.linear5.bias.detach()).contiguous()),
                    "dec1_w": c[0].weight.detach().reshape(512, FEAT_C),
                }
                # inference without the h1 / h2 outputs: linear2|3 -> linear4+5 -> decoder conv1 are three consecutive
                # linear maps (no activation in between, FaceRecon.py:151-165) and collapse into one (512, 1024) matrix
                w1d = c[0].weight.detach().reshape(512, FEAT_C).double()
                w45d, w23d = self._packs["ph_l45"][0].double(), self._packs["ph_l23"][0].double()
                b45d, b23d = self._packs["ph_l45"][1].double(), self._packs["ph_l23"][1].double()
                wf = w1d @ w45d
                self._packs["ph_fold"] = ((wf @ w23d).float().contiguous(), (wf @ b23d + w1d @ b45d).float().contiguous())
                # conv2 of the three pose tails as ONE grouped contraction (each alone: 257 row tiles on 148 SMs)
                self._packs["tails2"] = _Packed([g.conv2.weight, r.conv2.weight, t.conv2.weight],
                                                [_fold(g.conv2, g.bn2), _fold(r.conv2, r.bn2), _fold(t.conv2, t.bn2)], [0.0, 0.0, 0.0])
                # FACTORED first layers.  `feat` is a concatenation whose blocks fm_2 | fm_3 and fm_4 are nearest-neighbour
                # upsamplings of the 257 / 64 coarse points (FaceRecon.py:69-81), and a 1x1 convolution commutes with a
                # gather of rows:  W.[x | up(y)] = W_x.x + up(W_y.y).  So the 1024 upsampled input channels of the five
                # first-layer convolutions (the four of stage 1 + decoder conv1) are contracted ONCE PER COARSE POINT
                # (K = 512 at 257 and at 64 points per cloud instead of K = 1024 at 1028 points: 3.0x fewer flops for these
                # layers, which were 70 % of the network's), and the per-point contraction keeps K = 265 and adds the two
                # coarse products as gathered residuals in its epilogue (tgp_gemm_args.res1_idx / res2_idx).
                s1, d1 = self._packs["stage1"], self._packs["dec1"]
                self._packs["stage1f"] = _Packed.from_matrix(s1.w[:, _FINE_COLS], s1.scale, s1.shift, s1.slope)
                self._packs["dec1f"] = _Packed.from_matrix(d1.w[:, _FINE_COLS], d1.scale, d1.shift, d1.slope)
                self._packs["coarse1"] = _Packed.from_matrix(torch.cat([s1.w[:, _L1_COLS], d1.w[:, _L1_COLS]], 0))
                self._packs["coarse2"] = _Packed.from_matrix(torch.cat([s1.w[:, _L2_COLS], d1.w[:, _L2_COLS]], 0))
                for name, head in (("green", g), ("red", r), ("ts", t)):
                    sc, sh = _fold(head.conv3, head.bn3)
                    self._packs[name + "3"] = (head.conv3.weight.detach().reshape(256, 256), sc.contiguous(), sh.contiguous())
                    self._packs[name + "4"] = (head.conv4.weight.detach().reshape(head.k, 256), head.conv4.bias.detach())
            self._packs_key = key
        return self._packs

    @staticmethod
    def _stage(pk, x_split, K, outs, M, rows_per_group=0, shared_split=False, **kw):
        """one fused GEMM: column blocks of pk.w go to `outs` = [(n_cols, 'raw'|'split'|'max')]; returns the tensors.
        'max': per-cloud column max (torch.max over the points) taken in the epilogue, nothing else is written.
        shared_split: the 'split' outputs (all of one width n) are written side by side into ONE mixed operand of width
        n_split * n, the A operand of a grouped contraction (tgp_gemm_args.a_group_cols); that operand is returned for each."""
        segs, res, c0 = [], [], 0
        dev = x_split.device
        shared, n_sh, j_sh = None, sum(1 for _, kd in outs if kd == "split"), 0
        for n, kind in outs:
            if kind == "split" and shared_split:
                if shared is None:
                    shared = ops.mixed_buf(M, n_sh * n, dev)
                # column block j of the shared operand: 16-bit slot offset j*n = float offset j*n/2; slab width = operand Kp
                segs.append((c0, c0 + n, shared[:, j_sh * n // 2:], 5, ops.mixed_kpad(n_sh * n)))
                j_sh += 1
                res.append(shared)
                c0 += n
                continue
            if kind == "raw":
                t_ = torch.empty((M, n), dtype=torch.float32, device=dev)
                segs.append((c0, c0 + n, t_, 0, 0))
            elif kind == "max":
                t_ = torch.full((M // rows_per_group, n), -2 ** 31, dtype=torch.int32, device=dev)
                segs.append((c0, c0 + n, t_, 3, 0))
            else:
                t_ = ops.mixed_buf(M, n, dev)
                segs.append((c0, c0 + n, t_, 5, ops.mixed_kpad(n)))      # mode 5: mixed operand without the bf16(x) slot
            res.append(t_)
            c0 += n
        ops.gemm(None, pk.w, True, segs, scale=pk.scale, shift=pk.shift, neg_slope=pk.slope, K=K,
                 A_split=x_split, B_split=pk.w_split, rows_per_group=rows_per_group, mixed=2, **kw)
        return res

    def _forward_fused_eval(self, points, obj_id, enable_proj=False):
        mean = points.mean(dim=1, keepdim=True)
        centred = points - mean
        fa, g, r, t = self.face_all, self.rot_green, self.rot_red, self.ts
        enc = fa.encoder
        B, N, _ = centred.shape
        M = B * N
        parts = enc.encode(centred)
        pk = self._head_packs()
        kin = FEAT_C + 3
        # [feat | xyz] (Pose_Ts input, PoseNet9D.py:63) assembled straight into the tensor-core operand: upsampling
        # gathers, one-hot broadcast and both torch.cat of the reference in one launch
        one_hot = enc.one_hot(obj_id, B)
        flat = lambda t_: t_.reshape(-1, t_.shape[-1])
        raw = None
        if self.train_outputs:      # the (B, N, 1286) `feat` output itself (returned only with FLAGS.train, PoseNet9D.py:68-80)
            raw = ops.concat_rows(enc.concat_sources(parts, one_hot), B, N)[0]
        factored = self.factored_heads and parts["fm_2"].shape[1] * B >= 256 and parts["fm_4"].shape[1] * B >= 256
        if factored:
            # per-point operand [fm_0 | fm_1 | one-hot | xyz] (K = 265) and the two coarse operands, as mixed tensor-core operands
            N1, N2 = parts["fm_2"].shape[1], parts["fm_4"].shape[1]
            xs = ops.concat_rows([(flat(parts["fm_0"]), None, 1), (flat(parts["fm_1"]), None, 1), (one_hot, None, 0),
                                  (centred.reshape(M, 3), None, 1)], B, N, want_raw=False, want_split=True, mixed=True)[1]
            xs1 = ops.concat_rows([(flat(parts["fm_2"]), None, 1), (flat(parts["fm_3"]), None, 1)], B, N1, want_raw=False,
                                  want_split=True, mixed=True)[1]
            xs2 = ops.split_mixed(flat(parts["fm_4"]))
            # coarse products of all five first layers at once: (B*N1, 4608) and (B*N2, 4608), no bias / activation
            c1, c2 = pk["coarse1"], pk["coarse2"]
            nc = c1.w.shape[0]
            P1 = torch.empty((B * N1, nc), dtype=torch.float32, device=xs.device)
            P2 = torch.empty((B * N2, nc), dtype=torch.float32, device=xs.device)
            ops.gemm(None, c1.w, True, [(0, nc, P1, 0, 0)], K=c1.w.shape[1], A_split=xs1, B_split=c1.w_split, mixed=2, algo_flops=0)
            ops.gemm(None, c2.w, True, [(0, nc, P2, 0, 0)], K=c2.w.shape[1], A_split=xs2, B_split=c2.w_split, mixed=2, algo_flops=0)
            # rows of P1 / P2 that each level-0 point adds: its nearest coarse point (FaceRecon.py:69-73), as global row numbers
            gi1, gi2 = parts.get("up_rows") or enc.upsample_rows(parts["nn1"], parts["nn2"], N1, N2)
            k1 = len(_FINE_COLS)
            pk_s1, pk_d1 = pk["stage1f"], pk["dec1f"]
            n_s1, n_d1 = pk_s1.w.shape[0], pk_d1.w.shape[0]          # 4096 stage-1 columns, then the decoder's 512
            res_s1 = dict(res1=P1[:, :n_s1], res2=P2[:, :n_s1], res1_idx=gi1, res2_idx=gi2, algo_flops=2 * M * kin * n_s1)
            res_d1 = dict(res1=P1[:, n_s1:], res2=P2[:, n_s1:], res1_idx=gi1, res2_idx=gi2, algo_flops=2 * M * kin * n_d1)
        else:
            # [feat | xyz] (Pose_Ts input, PoseNet9D.py:63) assembled straight into the tensor-core operand: upsampling
            # gathers, one-hot broadcast and both torch.cat of the reference in one launch
            src = enc.concat_sources(parts, one_hot, extra=[(centred.reshape(M, 3), None, 1)])
            xs = ops.concat_rows(src, B, N, want_raw=False, want_split=True, mixed=True)[1]
            k1, res_s1, res_d1, pk_s1, pk_d1 = kin, {}, {}, pk["stage1"], pk["dec1"]
        # stage 1: four 1286/1289 -> 1024 convolutions as one contraction over the shared operand; conv_5's output is
        # only ever max-pooled over the cloud (FaceRecon.py:145-146), so that pooling happens in the epilogue
        # (the three 1024-wide hidden activations of the pose tails land side by side in one operand: their conv2 layers
        # run as one grouped contraction below)
        hid, _, f5max, _ = self._stage(pk_s1, xs, k1, [(1024, "split"), (1024, "split"), (1024, "max"),
                                                       (1024, "split")], M, rows_per_group=N, shared_split=True, **res_s1)
        # The three pose tails and the decoder chain are independent after stage 1.  Each of their GEMMs has 257 row tiles
        # for 148 SMs (a 1.7-wave tail), so they are issued on forked streams: the persistent CTAs of one kernel that finish
        # early free their SMs for the next kernel's CTAs.  Under CUDA-graph capture the forks become parallel branches.
        main = torch.cuda.current_stream()
        side = None
        if self.overlap_heads:
            if self._side is None:
                self._side = [torch.cuda.Stream(device=points.device) for _ in range(3)]
            side = self._side
            for st in side:
                st.wait_stream(main)
        pooled = ops.decode_max(f5max)
        # PH_Predictor tail (FaceRecon.py:147-165): per-cloud (M = batch) contractions on the skinny kernel
        w1, sc5, sh5, sl5 = pk["ph_l1"]
        feat_all = ops.linear_nk(pooled, w1, scale=sc5, shift=sh5, neg_slope=sl5, tc=False)      # linear1 + bn5 + leaky; dropout = id
        h1 = h2 = None
        dec = fa.decoder
        # Face_Dec on feat + cvec: W.(feat + c) = W.feat + W.c  -> per-cloud bias `gb` in the epilogue
        if self.train_outputs:
            w23, b23 = pk["ph_l23"]
            pi12 = ops.linear_nk(feat_all, w23, bias=b23, tc=False)                               # [pi1 | pi2]
            w45, b45 = pk["ph_l45"]
            cvec = ops.linear_nk(pi12, w45, bias=b45, tc=False)                                   # linear4(pi1) + linear5(pi2): (B,1286)
            oc = fa.ph_pred.output_channels
            h1, h2 = torch.sigmoid(pi12[:, :oc]), torch.sigmoid(pi12[:, oc:])
            gb = ops.linear_nk(cvec, pk["dec1_w"], tc=False)
        else:
            wfold, bfold = pk["ph_fold"]
            gb = ops.linear_nk(feat_all, wfold, bias=bfold, tc=False)
        (d1,) = self._stage(pk_d1, xs, k1, [(512, "split")], M, group_bias=gb, rows_per_group=N, **res_d1)
        (d2,) = self._stage(pk["dec2"], d1, 512, [(512, "split")], M)
        (d3,) = self._stage(pk["dec3"], d2, 512, [(256, "split")], M)
        (d4,) = self._stage(pk["dec4"], d3, 256, [(128, "raw")], M)
        last = dec.recon_head[3]
        recon = ops.linear_nk(d4, last.weight.reshape(3, 128), bias=last.bias).view(B, N, 3)

        def tails_pool():
            # conv2 + bn2 + relu of the three tails with the max over the points in the epilogue (PoseR.py:32-33): one grouped launch
            Bc = M // N
            hm = torch.full((Bc, 768), -2 ** 31, dtype=torch.int32, device=hid.device)
            p2 = pk["tails2"]
            ops.gemm(None, p2.w, True, [(0, 768, hm, 3, 0)], scale=p2.scale, shift=p2.shift, neg_slope=p2.slope, K=1024,
                     A_split=hid, B_split=p2.w_split, rows_per_group=N, mixed=2, a_kp=ops.mixed_kpad(3072), a_group_cols=256)
            return ops.decode_max(hm)

        def tail_fc(pooled3, j, name):
            # conv3 + bn3 + relu and conv4 of one tail on the per-cloud skinny kernel
            w3, sc3, sh3 = pk[name + "3"]
            v = ops.linear_nk(pooled3[:, 256 * j:256 * (j + 1)], w3, scale=sc3, shift=sh3, relu=True, tc=False)
            w4, b4 = pk[name + "4"]
            return ops.linear_nk(v, w4, bias=b4, tc=False)

        self._last_recon = recon          # (tests) the reference computes recon in inference too but only returns it with FLAGS.train
        names = ("green", "red", "ts")
        if side is not None:
            with torch.cuda.stream(side[0]):
                pooled3 = tails_pool()
            side[1].wait_stream(side[0])
            side[2].wait_stream(side[0])
            vecs = []
            for j, name in enumerate(names):      # the three short per-cloud chains run side by side
                with torch.cuda.stream(side[j]):
                    vecs.append(tail_fc(pooled3, j, name))
            for st in side:
                main.wait_stream(st)
        else:
            pooled3 = tails_pool()
            vecs = [tail_fc(pooled3, j, name) for j, name in enumerate(names)]
        green_R_vec, red_R_vec, ts_vec = vecs
        feat = feat_global = None
        if self.train_outputs:
            feat = raw.view(B, N, FEAT_C)
            feat_global = feat.max(dim=1)[0]
        return self._assemble(green_R_vec, red_R_vec, ts_vec[:, 0:3], ts_vec[:, 3:6], mean, recon, h1, h2, feat, feat_global)

    def _assemble(self, green_R_vec, red_R_vec, T, s, mean, recon, h1, h2, feat, feat_global):
        p_green_R = green_R_vec[:, 1:] / (torch.norm(green_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        p_red_R = red_R_vec[:, 1:] / (torch.norm(red_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        out = {'p_green_R': p_green_R, 'p_red_R': p_red_R, 'f_green_R': torch.sigmoid(green_R_vec[:, 0]),
               'f_red_R': torch.sigmoid(red_R_vec[:, 0]), 'Pred_T': T + mean.squeeze(1), 'Pred_s': s}
        if self.train_outputs:
            out.update({'recon': recon + mean, 'h1': h1, 'h2': h2, 'feat': feat, 'feat_global': feat_global})
        return out

    def forward(self, points, obj_id, enable_proj=False):
        _require_cuda(points, "PoseNet9D")
        # the fused inference pipeline bypasses autograd and the optional projection layer: take it only when nothing
        # can ask for a gradient and enable_proj is off (FaceRecon.py:83-84); otherwise the module-by-module path runs
        if (not self.training and not self.only_encoder and not enable_proj
                and not (torch.is_grad_enabled() and (points.requires_grad or any(p.requires_grad for p in self.parameters())))):
            return self._forward_fused_eval(points, obj_id, enable_proj)
        mean = points.mean(dim=1, keepdim=True)
        centred = points - mean
        if self.only_encoder:
            recon, _, feat_global_aug, _, _ = self.face_enc(centred, obj_id, enable_proj=enable_proj, pred_PH=False)
            return {'feat_global': feat_global_aug.max(2)[0], 'recon': recon}
        recon, feat, feat_global, h1, h2 = self.face_all(centred, obj_id, enable_proj=enable_proj, pred_PH=True)
        feat_global = feat_global.max(2)[0]
        feat_cf = feat.permute(0, 2, 1)
        green_R_vec = self.rot_green(feat_cf)
        red_R_vec = self.rot_red(feat_cf)
        p_green_R = green_R_vec[:, 1:] / (torch.norm(green_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        p_red_R = red_R_vec[:, 1:] / (torch.norm(red_R_vec[:, 1:], dim=1, keepdim=True) + 1e-6)
        f_green_R = torch.sigmoid(green_R_vec[:, 0])
        f_red_R = torch.sigmoid(red_R_vec[:, 0])
        feat_for_ts = torch.cat([feat, centred], dim=2)
        T, s = self.ts(feat_for_ts.permute(0, 2, 1))
        out = {'p_green_R': p_green_R, 'p_red_R': p_red_R, 'f_green_R': f_green_R, 'f_red_R': f_red_R,
               'Pred_T': T + mean.squeeze(1), 'Pred_s': s}
        if self.train_outputs:
            out.update({'recon': recon + mean, 'h1': h1, 'h2': h2, 'feat': feat, 'feat_global': feat_global})
        return out
