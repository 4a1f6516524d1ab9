"""Oracle: part-segmentation forward (test infrastructure, see oracle/__init__.py).

Follows part_segmentation/models/pt_mamba.py:631-787 (method HLT / SAST, eval mode) and
part_segmentation/models/pointnet2_utils.py:262-311 (PointNetFeaturePropagation, full sort as the reference).
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import mamba, model, spectral, tokenizer


def _bn(sd, prefix, x, eps=1e-5):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                        sd[prefix + "bias"], False, 0.0, eps)


def square_distance(src, dst):
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1)[:, :, None]
    dist += torch.sum(dst ** 2, -1)[:, None, :]
    return dist


def feature_propagation(sd, prefix, xyz1, xyz2, points1, points2):
    """xyz1 (B,N,3), xyz2 (B,S,3), points1 (B,N,D1), points2 (B,S,D2) -> (B, D', N)."""
    B, N, _ = xyz1.shape
    dists, idx = square_distance(xyz1, xyz2).sort(dim=-1)
    dists, idx = dists[:, :, :3], idx[:, :, :3]
    recip = 1.0 / (dists + 1e-8)
    weight = recip / recip.sum(dim=2, keepdim=True)
    gathered = torch.gather(points2[:, None].expand(-1, N, -1, -1), 2, idx[..., None].expand(-1, -1, -1, points2.shape[-1]))
    interp = (gathered * weight[..., None]).sum(dim=2)
    x = torch.cat([points1, interp], dim=-1).permute(0, 2, 1)
    i = 0
    while (prefix + f"mlp_convs.{i}.weight") in sd:
        x = F.conv1d(x, sd[prefix + f"mlp_convs.{i}.weight"], sd[prefix + f"mlp_convs.{i}.bias"])
        x = F.relu(_bn(sd, prefix + f"mlp_bns.{i}.", x))
        i += 1
    return x


def seg_forward(sd: dict, cfg: dict, pts: torch.Tensor, cls_label: torch.Tensor, noise: torch.Tensor,
                order_override=None, perm_override=None):
    """get_model.forward(pts (B,3,N), cls_label (B,16)) -> log-probs (B,N,cls_dim); ``noise`` is the U[0,1) HLT
    tie-break the reference draws with torch.rand (pt_mamba.py:673)."""
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    B, _, N = pts.shape
    p = pts.transpose(1, 2).contiguous()
    nbr, center, org, fidx, kidx = tokenizer.group(p, 128, 32)
    tok = model.encoder(sd, "encoder.", nbr)
    pos = model.pos_embed(sd, "pos_embed.", center)
    k = cfg["k_top_eigenvectors"]
    vals, vecs, allv, S = spectral.spectral_eig(center, cfg["knn_graph"], cfg["alpha"], cfg["symmetric"],
                                                cfg["self_loop"], cfg["binary"], k, cfg["smallest"])
    if cfg["method"] == "HLT":
        order = spectral.hlt_order(vecs.float(), k, noise) if order_override is None else order_override
        x = spectral.hlt_layout(tok, order, k, cfg["reverse"])
        sp = spectral.hlt_layout(pos, order, k, cfg["reverse"])
        sc = spectral.hlt_layout(center, order, k, cfg["reverse"])
    else:
        perm = spectral.sast_perm(vecs) if perm_override is None else perm_override
        order = perm
        x = spectral.order_gather(tok, perm, cfg["reverse"])
        sp = spectral.order_gather(pos, perm, cfg["reverse"])
        sc = spectral.order_gather(center, perm, cfg["reverse"])
    feats = mamba.mixer_model(sd, "blocks.", x, sp, cfg["depth"], fetch_idx=cfg["fetch_idx"])
    feats = [mamba.layer_norm(sd, "norm.", f).transpose(1, 2) for f in feats]
    xf = torch.cat(feats, dim=1)
    x_max = xf.max(dim=2)[0]
    x_avg = xf.mean(dim=2)
    lab = F.conv1d(cls_label.view(B, 16, 1), sd["label_conv.0.weight"])
    lab = F.leaky_relu(_bn(sd, "label_conv.1.", lab), 0.2)
    glob = torch.cat((x_max[..., None].expand(-1, -1, N), x_avg[..., None].expand(-1, -1, N), lab.expand(-1, -1, N)), 1)
    f0 = feature_propagation(sd, "propagation_0.", p, sc, p, xf.transpose(1, 2))
    h = torch.cat((f0, glob), 1)
    h = F.relu(_bn(sd, "bns1.", F.conv1d(h, sd["convs1.weight"], sd["convs1.bias"])))
    h = F.relu(_bn(sd, "bns2.", F.conv1d(h, sd["convs2.weight"], sd["convs2.bias"])))
    h = F.conv1d(h, sd["convs3.weight"], sd["convs3.bias"])
    return F.log_softmax(h, dim=1).permute(0, 2, 1), dict(order=order, eigvecs=vecs)
