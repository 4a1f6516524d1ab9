"""MAE pre-training path (config C3): ``MaskMamba_2`` encoder, ``MambaDecoder_SST`` decoder and
``Point_MAE_Mamba`` with the reference's class names, config keys and state-dict keys
(models/point_mamba.py:2135-2541, 2837-2866, 2869-3219; cfgs/pretrain.yaml:35-65).

The reference HEAD only runs with ``use_wavelets=True`` (its spectral branch sets the eigenvectors to None,
SURVEY.md section 3 "HEAD caveats"); what is implemented here is the spectral path those lines were written
for - the batched graph/eigh of :2958-3050, the masked sort of MaskMamba_3.forward :2734-2796 and the token
restore of :3147-3197 - on the sm_100a kernels: one spectral kernel, then row-gather kernels driven by small
integer index maps (si_mamba_b200/layout.py) instead of ~60 boolean-mask / torch.where launches with host syncs.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import layout, ops
from .point_mamba import Encoder, Group, MixerModel


def chamfer_l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """pytorch3d ``chamfer_distance(x, y, batch_reduction=None)[0]`` with squared L2 and point_reduction="mean"
    (models/point_mamba.py:2950, 3203): (N,P,3), (N,Q,3) -> (N,).  One warp per pair on the device
    (csrc/chamfer.cu, sim_chamfer_l2_fwd / _bwd).  No CPU / eager fallback: CPU tensors or sets above 256 points raise."""
    if not x.is_cuda or x.shape[1] > 256 or y.shape[1] > 256:
        raise RuntimeError("chamfer_l2 runs on CUDA tensors with at most 256 points per set (there is no CPU fallback)")
    return ops.chamfer_l2(x, y)


def rand_mask_host(B: int, G: int, mask_ratio: float) -> torch.Tensor:
    """models/point_mamba.py:2232-2255 ``_mask_center_rand``: per cloud a numpy shuffle of G-m zeros and m ones,
    m = int(ratio * G).  Host tensor (B, G) bool; callers that capture the step in a CUDA graph copy it into a static
    device tensor and pass it as ``bool_masked_pos``."""
    m = int(mask_ratio * G)
    out = np.zeros([B, G])
    for i in range(B):
        mask = np.hstack([np.zeros(G - m), np.ones(m)])
        np.random.shuffle(mask)
        out[i, :] = mask
    return torch.from_numpy(out).to(torch.bool)


def _init_trunc_normal(m):
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv1d):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


class MaskMamba_2(nn.Module):
    """MAE encoder (models/point_mamba.py:2135-2541): mask, spectral sort of the visible tokens, 12 Mamba blocks."""

    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        tc = config.transformer_config
        self.mask_ratio = tc.mask_ratio
        self.group_size = config.group_size
        self.num_group = config.num_group
        self.trans_dim = tc.trans_dim
        self.depth = tc.depth
        self.num_heads = tc.num_heads
        self.k_top_eigenvectors = tc.k_top_eigenvectors
        self.encoder_dims = tc.encoder_dims
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.mask_type = tc.mask_type
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.trans_dim))
        self.blocks = MixerModel(d_model=self.trans_dim, n_layer=self.depth, rms_norm=self.config.rms_norm)
        self.norm = nn.LayerNorm(self.trans_dim)
        self.apply(_init_trunc_normal)

    def _mask_center_rand(self, center, noaug=False):
        """models/point_mamba.py:2232-2255: per cloud a numpy shuffle of G-m zeros and m ones, m = int(ratio * G)."""
        B, G, _ = center.shape
        if noaug or self.mask_ratio == 0:
            return torch.zeros(center.shape[:2], dtype=torch.bool, device=center.device)
        self.num_mask = int(self.mask_ratio * G)
        overall_mask = np.zeros([B, G])
        for i in range(B):
            mask = np.hstack([np.zeros(G - self.num_mask), np.ones(self.num_mask)])
            np.random.shuffle(mask)
            overall_mask[i, :] = mask
        return torch.from_numpy(overall_mask).to(torch.bool).to(center.device)

    def forward(self, neighborhood, center, perm, reverse=True, noaug=False, bool_masked_pos=None, n_vis=None):
        """-> (x_vis (B, 2k*n_vis, C), maps) with ``maps`` the index maps of layout.mae_index_maps."""
        if not reverse:
            raise NotImplementedError("the MAE path is only defined for reverse=True (point_mamba.py:2778-2796)")
        user_mask = bool_masked_pos is not None
        if bool_masked_pos is None:
            if self.mask_type != 'rand':
                raise NotImplementedError("mask_type 'block' is not used by cfgs/pretrain.yaml")
            bool_masked_pos = self._mask_center_rand(center, noaug=noaug)
        tokens = self.encoder(neighborhood)
        pos = self.pos_embed(center)
        G = center.shape[1]
        if user_mask:
            pass  # the caller's n_vis, or None = counted from the mask (host sync + validation)
        elif noaug or self.mask_ratio == 0:
            n_vis = G
        else:
            n_vis = G - int(self.mask_ratio * G)
        maps = layout.mae_index_maps(perm, bool_masked_pos, n_vis)
        # masked sort = row compaction through the index maps (sim_mae_compact_fwd / _bwd)
        x_vis = ops.MaeCompact.apply(tokens, maps["src_vis"], maps["inv_vis"])
        pos_vis = ops.MaeCompact.apply(pos, maps["src_vis"], maps["inv_vis"])
        x_vis = self.norm(self.blocks(x_vis, pos_vis))
        maps["pos"] = pos
        return x_vis, maps


class MambaDecoder_SST(nn.Module):
    """models/point_mamba.py:2837-2866."""

    def __init__(self, embed_dim=384, depth=4, norm_layer=nn.LayerNorm, config=None):
        super().__init__()
        self.blocks = MixerModel(d_model=embed_dim, n_layer=depth, rms_norm=config.rms_norm, drop_path=config.drop_path)
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x, pos, return_token_num=None):
        return self.head(self.norm(self.blocks(x, pos)))


class Point_MAE_Mamba(nn.Module):
    """models/point_mamba.py:2869-3219, method ``smallest_eigenvectors_seperate_learnable_tokens``."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        tc = config.transformer_config
        self.trans_dim = tc.trans_dim
        if tc.method != "smallest_eigenvectors_seperate_learnable_tokens":
            raise NotImplementedError(f"MAE method {tc.method!r}: only the spectral method is on the hot path")
        self.MAE_encoder = MaskMamba_2(config)
        self.group_size = config.group_size
        self.num_group = config.num_group
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.trans_dim))
        self.decoder_pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.trans_dim))
        self.decoder_depth = tc.decoder_depth
        self.MAE_decoder = MambaDecoder_SST(embed_dim=self.trans_dim, depth=self.decoder_depth, config=config)
        self.group_divider = Group(num_group=self.num_group, group_size=self.group_size)
        self.increase_dim = nn.Sequential(nn.Conv1d(self.trans_dim, 3 * self.group_size, 1))
        nn.init.trunc_normal_(self.mask_token, std=.02)
        self.loss = config.loss
        if self.loss != "cdl2":
            raise NotImplementedError("only the Chamfer-L2 loss of cfgs/pretrain.yaml is built")
        self.loss_func = chamfer_l2
        self.method = tc.method
        self.reverse = tc.reverse
        self.k_top_eigenvectors = tc.k_top_eigenvectors
        self.smallest = tc.smallest
        self.knn_graph = tc.knn_graph
        self.alpha = tc.alpha
        self.symmetric = tc.symmetric
        self.self_loop = tc.self_loop
        self.binary = tc.binary
        self.register_buffer('baseline', torch.tensor(torch.inf))

    def spectral_order(self, center):
        """Batched graph + Laplacian (deg.clamp(1e-12) variant, :3001-3050) + eigensolver + argsort, one kernel."""
        return ops.spectral_eig(center, self.knn_graph, self.alpha, self.symmetric, self.self_loop, self.binary,
                                self.k_top_eigenvectors, self.smallest, eps_mode="clamp1e-12")

    def forward(self, pts, noaug=False, vis=False, tau=None, use_wavelets: bool = False, use_diff_sort: bool = False,
                ret_policy: bool = False, ret_only_policy: bool = False, save_pts_dir: str = None, epoch: int = None,
                bool_masked_pos=None, n_vis=None, **kwargs):
        """pts (B,N,3) -> scalar Chamfer-L2 loss (x_vis when ``noaug``)."""
        if use_wavelets or use_diff_sort or ret_only_policy:
            raise NotImplementedError("wavelet / learned-ordering branches are out of the hot-path scope")
        neighborhood, center, neighborhood_org = self.group_divider(pts)
        perm = self.spectral_order(center)["perm"]
        x_vis, maps = self.MAE_encoder(neighborhood, center, perm, self.reverse, noaug, bool_masked_pos, n_vis)
        if noaug:
            return x_vis
        B, _, C = x_vis.shape
        # token restore: decoder position t shows the mask token or the encoder row with the same visible rank
        x_full = ops.MaeRestore.apply(x_vis, self.mask_token, maps["restore_src"], maps["vis_pos"])
        pos_full = layout.gather_rows(maps["pos"], maps["perm_full"], fanout=2 * perm.shape[1])
        x_rec = self.MAE_decoder(x_full, pos_full, None)
        x_rec = layout.gather_rows(x_rec, maps["rec_src"], fanout=1)                    # (B, 2k*m, C) masked positions
        M = x_rec.shape[1]
        rebuild_points = self.increase_dim(x_rec.transpose(1, 2)).transpose(1, 2).reshape(B * M, -1, 3)
        patch_of_rec = torch.gather(maps["perm_full"].long(), 1, maps["rec_src"].long())   # (B, 2k*m)
        gt_points = torch.gather(neighborhood, 1, patch_of_rec[..., None, None].expand(-1, -1, self.group_size, 3))
        loss = self.loss_func(rebuild_points.float(), gt_points.reshape(B * M, -1, 3).float()).mean()
        if ret_policy:
            return loss, torch.zeros(B, device=pts.device)
        return loss
