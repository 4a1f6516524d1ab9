"""Part-segmentation encoder path (config C4): ``get_model`` with the reference's constructor / forward signature,
config keys and state-dict keys (part_segmentation/models/pt_mamba.py:419-797), ``MixerModelForSegmentation``
(:228-416) and ``PointNetFeaturePropagation`` (part_segmentation/models/pointnet2_utils.py:262-311).

Encoder part on the sm_100a kernels: FPS/kNN, one spectral kernel, HLT bucket ids (multilevel_travers :595-607) +
the chunked forward/reverse layout (:670-723, including its overwrite quirk) as ONE row-gather
driven by an index map, then the Mamba stack tapping layers 3/7/11.  Feature propagation and the per-point head
stay PyTorch; the 3-NN uses top-k instead of the reference's full sort.
"""

from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layout, ops
from .block import DropPath, fused_add_norm
from .point_mamba import Encoder, Group, RMSNorm, _init_weights, create_block


def part_seg_config():
    """part_segmentation/cfgs/config.yaml:1-18."""
    from .config import Config
    return Config(model="pt_mamba", trans_dim=384, depth=12, drop_path_rate=0.1, rms_norm=False, use_cls_token=False,
                  drop_path=0.2, drop_out=0., fetch_idx=[3, 7, 11], method="HLT", reverse=True, k_top_eigenvectors=4,
                  smallest=True, knn_graph=10, symmetric=True, self_loop=True, alpha=10., binary=False)


class MixerModelForSegmentation(nn.Module):
    """pt_mamba.py:228-416: the Mamba stack returning norm_f(h + residual) at the layers in ``fetch_idx``."""

    def __init__(self, d_model: int, n_layer: int, ssm_cfg=None, norm_epsilon: float = 1e-5, rms_norm: bool = False,
                 initializer_cfg=None, fused_add_norm=False, residual_in_fp32=False, drop_out_in_block: int = 0.,
                 drop_path: int = 0.1, fetch_idx=(3, 7, 11), device=None, dtype=None) -> None:
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.layers = nn.ModuleList([
            create_block(d_model, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm,
                         residual_in_fp32=residual_in_fp32, fused_add_norm=fused_add_norm, layer_idx=i,
                         drop_path=drop_path, **factory_kwargs) for i in range(n_layer)])
        self.norm_f = (nn.LayerNorm if not rms_norm else RMSNorm)(d_model, eps=norm_epsilon, **factory_kwargs)
        self.apply(partial(_init_weights, n_layer=n_layer, **(initializer_cfg if initializer_cfg is not None else {})))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.fetch_idx = list(fetch_idx)

    def forward(self, input_ids, pos, inference_params=None):
        hidden_states = input_ids + pos
        residual = None
        feature_list = []
        for idx, layer in enumerate(self.layers):
            hidden_states, residual = layer(hidden_states, residual, inference_params=inference_params)
            if idx in self.fetch_idx:
                out, _ = fused_add_norm(self.norm_f, hidden_states, residual, want_residual=False)
                feature_list.append(out)
        return feature_list


class PointNetFeaturePropagation(nn.Module):
    """pointnet2_utils.py:262-311 (3-NN inverse-squared-distance interpolation + shared MLP)."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        xyz1 = xyz1.permute(0, 2, 1)
        xyz2 = xyz2.permute(0, 2, 1)
        points2 = points2.permute(0, 2, 1)
        B, N, C = xyz1.shape
        _, S, _ = xyz2.shape
        if S == 1:
            interpolated_points = points2.repeat(1, N, 1)
        else:
            if not xyz1.is_cuda or not 3 <= S <= 2048:
                raise RuntimeError("PointNetFeaturePropagation: CUDA tensors with 3 <= S <= 2048 centres only (no CPU fallback)")
            # one kernel: centres staged in shared memory, warp-level top-3, weighted row gather (sim_three_nn_interp_fwd)
            interpolated_points = ops.three_nn_interpolate(xyz1, xyz2, points2)
        if points1 is not None:
            new_points = torch.cat([points1.permute(0, 2, 1), interpolated_points], dim=-1)
        else:
            new_points = interpolated_points
        new_points = new_points.permute(0, 2, 1)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_points = F.relu(bn(conv(new_points)))
        return new_points


class SegGroup(Group):
    """pt_mamba.py:158-191: the segmentation flavour of Group returns (neighborhood, center)."""

    def forward(self, xyz):
        neighborhood, center, _ = super().forward(xyz)
        return neighborhood, center


class get_model(nn.Module):
    def __init__(self, cls_dim, config=None):
        super().__init__()
        self.trans_dim = config.trans_dim
        self.depth = config.depth
        self.cls_dim = cls_dim
        self.group_size = 32
        self.num_group = 128
        self.group_divider = SegGroup(num_group=self.num_group, group_size=self.group_size)
        self.encoder_dims = 384
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.trans_dim))
        self.blocks = MixerModelForSegmentation(d_model=self.trans_dim, n_layer=self.depth, rms_norm=config.rms_norm,
                                                drop_path=config.drop_path, fetch_idx=config.fetch_idx)
        self.drop_out = nn.Dropout(config.drop_out) if "drop_out" in config else nn.Dropout(0)
        self.drop_path_rate = config.drop_path_rate
        self.drop_path_block = DropPath(self.drop_path_rate) if self.drop_path_rate > 0. else nn.Identity()
        self.norm = nn.LayerNorm(self.trans_dim)
        self.label_conv = nn.Sequential(nn.Conv1d(16, 64, kernel_size=1, bias=False), nn.BatchNorm1d(64),
                                        nn.LeakyReLU(0.2))
        self.propagation_0 = PointNetFeaturePropagation(in_channel=3 * self.trans_dim + 3, mlp=[self.trans_dim * 4, 1024])
        self.convs1 = nn.Conv1d(2 * 3 * self.trans_dim + 64 + 1024, 512, 1)
        self.dp1 = nn.Dropout(0.5)
        self.convs2 = nn.Conv1d(512, 256, 1)
        self.convs3 = nn.Conv1d(256, self.cls_dim, 1)
        self.bns1 = nn.BatchNorm1d(512)
        self.bns2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()
        self.method = config.method
        self.reverse = config.reverse
        self.k_top_eigenvectors = config.k_top_eigenvectors
        self.smallest = config.smallest
        self.knn_graph = config.knn_graph
        self.symmetric = config.symmetric
        self.self_loop = config.self_loop
        self.alpha = config.alpha
        self.binary = config.binary
        self.loss_ce = nn.CrossEntropyLoss()

    def get_loss_acc(self, ret, gt):
        loss = self.loss_ce(ret, gt.long())
        pred = ret.argmax(-1)
        acc = (pred == gt).sum() / float(gt.size(0))
        return loss, acc * 100

    def multilevel_travers(self, eigen_vectors, level):
        """pt_mamba.py:595-607: sign bits against the per-vector mean -> bucket id."""
        means = eigen_vectors.mean(dim=1, keepdim=True)
        binaries = (eigen_vectors >= means)[:, :, :level]
        powers_of_2 = 2 ** torch.arange(start=level - 1, end=-1, step=-1, device=eigen_vectors.device)
        return torch.sum(binaries * powers_of_2[None, None, :], dim=-1)

    def forward(self, pts, cls_label, hlt_noise=None):
        """pts (B,3,N), cls_label (B,16) one-hot -> log-probabilities (B,N,cls_dim)."""
        B, C, N = pts.shape
        pts = pts.transpose(-1, -2).contiguous()
        neighborhood, center = self.group_divider(pts)
        group_input_tokens = self.encoder(neighborhood)
        pos = self.pos_embed(center)
        spec = ops.spectral_eig(center, self.knn_graph, self.alpha, self.symmetric, self.self_loop, self.binary,
                                self.k_top_eigenvectors, self.smallest)
        if self.method == "HLT":
            ids = self.multilevel_travers(spec["vecs"], self.k_top_eigenvectors).float()
            if hlt_noise is None:
                hlt_noise = torch.rand(ids.shape[0], ids.shape[1])  # CPU RNG then moved, as the reference (:673)
            keys = ids + hlt_noise.to(ids.device)
            order, _ = ops.argsort_rows(keys.contiguous())
            src = layout.hlt_src_index(order, self.k_top_eigenvectors, bool(self.reverse))      # (B, 2G), -1 = zero
            x = layout.gather_rows(group_input_tokens, src, fanout=2)
            sorted_pos = layout.gather_rows(pos, src, fanout=2)
            valid = (src >= 0)[..., None]
            sorted_center = torch.gather(center, 1, src.clamp(min=0).long()[..., None].expand(-1, -1, 3)) * valid
        elif self.method == "SAST":
            perm, inv = spec["perm"], spec["inv_perm"]
            x = ops.order_gather(group_input_tokens, perm, bool(self.reverse), inv)
            sorted_pos = ops.order_gather(pos, perm, bool(self.reverse), inv)
            flat = perm.reshape(B, -1).long()
            if self.reverse:
                flat = torch.cat((flat, flat.flip(1)), dim=1)
            sorted_center = torch.gather(center, 1, flat[..., None].expand(-1, -1, 3))
        else:
            raise NotImplementedError(f"method {self.method!r}")
        feature_list = self.blocks(x, sorted_pos)
        feature_list = [self.norm(f.float()).transpose(-1, -2).contiguous() for f in feature_list]
        x = torch.cat(feature_list, dim=1)
        x_max = torch.max(x, 2)[0]
        x_avg = torch.mean(x, 2)
        x_max_feature = x_max.view(B, -1).unsqueeze(-1).repeat(1, 1, N)
        x_avg_feature = x_avg.view(B, -1).unsqueeze(-1).repeat(1, 1, N)
        cls_label_feature = self.label_conv(cls_label.view(B, 16, 1)).repeat(1, 1, N)
        x_global_feature = torch.cat((x_max_feature, x_avg_feature, cls_label_feature), 1)
        f_level_0 = self.propagation_0(pts.transpose(-1, -2), sorted_center.transpose(-1, -2), pts.transpose(-1, -2), x)
        x = torch.cat((f_level_0, x_global_feature), 1)
        x = self.relu(self.bns1(self.convs1(x)))
        x = self.dp1(x)
        x = self.relu(self.bns2(self.convs2(x)))
        x = self.convs3(x)
        x = F.log_softmax(x, dim=1)
        return x.permute(0, 2, 1)
