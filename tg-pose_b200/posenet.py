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


def pointwise(x_cl, conv, bn=None, act=None, slope=0.2):
    """1x1 Conv1d (+ eval BatchNorm + activation) on a channel-last (B,N,Cin) tensor -> (B,N,Cout),
    one GEMM launch with the affine/ReLU folded into the epilogue.  LeakyReLU is applied afterwards."""
    B, N, cin = x_cl.shape
    w = conv.weight.reshape(conv.out_channels, cin)
    scale = shift = None
    if bn is not None:
        scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        if conv.bias is not None:
            shift = shift + conv.bias * scale
        bias = None
    else:
        bias = conv.bias
    out = ops.linear_nk(x_cl.reshape(B * N, cin), w, bias=bias, scale=scale, shift=shift, relu=(act == "relu"))
    out = out.view(B, N, conv.out_channels)
    if act == "leaky":
        out = F.leaky_relu(out, slope, inplace=True)
    return out


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
        if self.training:
            return self.recon_head(self.conv1d_block(x)).permute(0, 2, 1)
        h = x.permute(0, 2, 1).contiguous()
        b = self.conv1d_block
        h = pointwise(h, b[0], b[1], "relu")
        h = pointwise(h, b[3], b[4], "relu")
        h = pointwise(h, b[6], b[7], "relu")
        r = self.recon_head
        h = pointwise(h, r[0], r[1], "relu")
        return pointwise(h, r[3])


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
        if self.training:
            f = self.conv_5(feat.permute(0, 2, 1))
            pooled = F.adaptive_max_pool1d(f, 1).view(bs, -1)
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
        if self.training:
            x = F.relu(self.bn1(self.conv1(x)))
            x = F.relu(self.bn2(self.conv2(x)))
            x = torch.max(x, 2, keepdim=True)[0]
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
        if not only_encoder:
            self.face_all = FaceNet(**enc_kwargs)
            self.rot_green = Rot_green()
            self.rot_red = Rot_red()
            self.ts = Pose_Ts()
        else:
            self.face_enc = FaceNet(**enc_kwargs)

    def forward(self, points, obj_id, enable_proj=False):
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
